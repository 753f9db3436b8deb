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


def _dir_schedule(count, numframe):
    """src/main.cpp:712-731: source pair and timestep of output frame i"""
    import math
    scale = count / numframe
    out = []
    for i in range(numframe):
        fx = np.float32(i * scale)
        sx = int(math.floor(fx))
        fx = np.float32(fx - sx)
        if sx >= count - 1:
            sx, fx = count - 2, np.float32(1.0)
        out.append((sx, float(fx)))
    return out


@pytest.mark.parametrize("gpus,jobs", [("0", "1:2:2"), ("0,1", "1:2,2:2")])
def test_reference_cli_directory_mode(tmp_path, gpus, jobs):
    """Directory mode, the way a video is interpolated (`-i frames/ -o out/`): the reference's load / proc / save threads drive
    ONE RIFE object per `-g` id from several proc threads each (src/main.cpp:819-866, 346-366).  `-g 0,1` is the reference's
    in-process multi-GPU mode: one process, one engine per device.  Every written frame is compared with the oracle (or with the
    source frame for timestep 0 / 1)."""
    if not os.path.exists(CLI):
        pytest.skip("host/_cli/rife-b200-cli not built (needs /root/reference at build time)")
    md = os.path.join(parity.REF_DIR, "models", "rife-v4.6")
    if not os.path.isdir(md):
        pytest.skip("reference model files not shipped")
    import __graft_entry__ as g
    ndev = g.load_package().device_count()
    if len(gpus.split(",")) > ndev:
        pytest.skip("needs %d CUDA devices, %d visible" % (len(gpus.split(",")), ndev))
    from PIL import Image
    w, h, count = 320, 192, 5
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir()
    outd.mkdir()
    frames = [parity.synth.frame(k, w, h) for k in range(count)]
    for k, f in enumerate(frames):
        Image.fromarray(f).save(str(ind / ("%08d.png" % (k + 1))))
    env = dict(os.environ, RIFE_B200_LIB=LIB)
    r = subprocess.run([CLI, "-i", str(ind), "-o", str(outd), "-m", md, "-g", gpus, "-j", jobs], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    sched = _dir_schedule(count, 2 * count)
    for i, (sx, t) in enumerate(sched):
        p = outd / ("%08d.png" % (i + 1))
        assert p.exists(), (i, r.stderr[-1000:])
        got = np.array(Image.open(str(p)).convert("RGB"))
        if t == 0.0:
            assert np.array_equal(got, frames[sx]), i
        elif t == 1.0:
            assert np.array_equal(got, frames[sx + 1]), i
        else:
            ref, _ = parity.run_oracle("rife-v4.6", frames[sx], frames[sx + 1], t)
            res = parity.compare(got, ref)
            assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, (i, sx, t, res)
