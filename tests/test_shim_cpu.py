"""Host shim (class RIFE + ncnn-namespace compat headers): builds, refuses CPU mode, and the reference's own
src/main.cpp compiles unchanged against it (syntax check; only where /root/reference is mounted)."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "host")
LIBDIR = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "lib")


def test_shim_probe_runs_without_gpu():
    exe = os.path.join(LIBDIR, "shim_demo")
    if not os.path.exists(exe):
        pytest.skip("host shim not built")
    env = dict(os.environ, RIFE_B200_LIB=os.path.join(LIBDIR, "librife_b200.so"))
    r = subprocess.run([exe, "--probe"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert '"mat_ok": 1' in r.stdout
    assert "no CPU path" in r.stderr


def test_reference_main_cpp_compiles_against_shim():
    ref = "/root/reference/src"
    if not os.path.exists(os.path.join(ref, "main.cpp")):
        pytest.skip("reference not mounted")
    with tempfile.TemporaryDirectory() as td:
        shutil.copy(os.path.join(ref, "main.cpp"), td)  # quoted includes resolve next to the file first
        cmd = ["g++", "-std=c++11", "-fsyntax-only", "-fopenmp", "-I" + HOST, "-I" + os.path.join(HOST, "ncnn_compat"), "-I" + ref,
               "-I" + os.path.join(ref, "libwebp", "src"), os.path.join(td, "main.cpp")]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
