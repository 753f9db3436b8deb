"""Error histogram of the fused path vs the oracle for different precision choices (which IFBlocks run their residual
chain on plain fp16 activations).  Prints one JSON line per (case, mask)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import parity
pkg = g.load_package()

def real_pair():
    try:
        from PIL import Image
        d = os.path.join(ROOT, "oracle", "_ref", "images")
        a = np.array(Image.open(os.path.join(d, "0.png")).convert("RGB"))
        b = np.array(Image.open(os.path.join(d, "1.png")).convert("RGB"))
        return a, b
    except Exception as e:
        print("no real images:", e)
        return None

cases = [("synth_640x360", parity.synth.pair(640, 360)), ("synth_large_motion_640x352", parity.synth.pair(640, 352, dx=24, dy=16)),
         ("synth_1920x1080", parity.synth.pair(1920, 1080)), ("synth_1920x1080_fast_motion", parity.synth.pair(1920, 1080, dx=17, dy=9, seed=3))]
rp = real_pair()
if rp is not None:
    cases.append(("readme_images_640x360", rp))
r = pkg.RIFE(0, False, False, False, 1, False, True)
r.load(parity.model_dir("rife-v4.6"))
for name, (a, b) in cases:
    ref, _ = parity.run_oracle("rife-v4.6", a, b, 0.5, threads=16)
    for mask in [int(m, 0) for m in os.environ.get("RIFE_STUDY_MASKS", "0,8,12,14,15").split(",")]:
        r.set_option("plain_blocks", mask)
        out = r.process(a, b, 0.5)
        res = parity.compare(out, ref)
        res.update(case=name, plain_blocks=mask)
        print(json.dumps(res))
r.close()
