"""The reference's UNMODIFIED src/main.cpp, linked against this repo's `class RIFE` shim (host/build_cli.py), run as the
reference CLI would be: `-0 frame0.png -1 frame1.png -o out.png -m <modeldir> -g 0`, compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "host", "_cli", "rife-b200-cli")
LIB = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "lib", "librife_b200.so")


@pytest.mark.parametrize("model,extra", [("rife-v4.6", []), ("rife-v4.6", ["-s", "0.25"]), ("rife-anime", [])])
def test_reference_cli_runs_on_the_b200_engine(tmp_path, model, extra):
    if not os.path.exists(CLI):
        pytest.skip("host/_cli/rife-b200-cli not built (needs /root/reference at build time)")
    md = os.path.join(parity.REF_DIR, "models", model)
    f0 = os.path.join(parity.REF_DIR, "images", "0.png")
    f1 = os.path.join(parity.REF_DIR, "images", "1.png")
    if not (os.path.isdir(md) and os.path.exists(f0)):
        pytest.skip("reference model / README frames not shipped")
    from PIL import Image
    out = str(tmp_path / "out.png")
    env = dict(os.environ, RIFE_B200_LIB=LIB)
    r = subprocess.run([CLI, "-0", f0, "-1", f1, "-o", out, "-m", md, "-g", "0"] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert r.returncode == 0 and os.path.exists(out), r.stderr[-2000:]
    got = np.array(Image.open(out).convert("RGB"))
    a = np.array(Image.open(f0).convert("RGB"))
    b = np.array(Image.open(f1).convert("RGB"))
    t = float(extra[1]) if extra else 0.5
    ref, _ = parity.run_oracle(model, a, b, t)
    res = parity.compare(got, ref)
    assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, res
