"""Host emulation of the fused rife-v4.6 HBM kernels (tests/emu/emu_fused.cpp): the device code of
fused_v46_kernels.cuh compiled with g++ (grid = loops), checked (a) across the three flow / mask storage variants
(recompute_fm 0 / 1 / 2: bit-identical head tensors and frames) and (b) against a whole-image restatement of
SURVEY.md Appendix B.  Needs only g++ and the CUDA headers (cuda_fp16.h, vector_types.h); no GPU, no CUDA runtime."""
import ctypes
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_include():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_fp16.h")):
            return os.path.join(d, "include")
    return None


def _build_and_run(tmp_path, lean):
    inc = _cuda_include()
    if shutil.which("g++") is None or inc is None:
        pytest.skip("g++ or CUDA headers not available")
    exe = str(tmp_path / ("emu_fused_lean%d" % lean))
    src = os.path.join(ROOT, "tests", "emu", "emu_fused.cpp")
    csrc = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "csrc")
    # -ffp-contract=off: every variant must round identically; the comparison is bit-exact
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-DRIFE_FUSED_LEAN=%d" % lean, "-I" + inc, "-I" + csrc, src, "-o", exe],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "EMU OK" in r.stdout, r.stdout[-4000:]
    return [l for l in r.stdout.splitlines() if l.startswith("checksum")][0]


def test_fused_kernels_host_emulation(tmp_path):
    _build_and_run(tmp_path, 0)


def test_lean_build_of_the_fused_kernels_is_bit_identical(tmp_path):
    """RIFE_FUSED_LEAN=1 (integer lin_coeff, packed half conversions): same head tensors and frames, byte for byte."""
    assert _build_and_run(tmp_path, 0) == _build_and_run(tmp_path, 1)


# ---- csrc/hbm_kernels.cu on the host (tests/emu/emu_hbm.cpp) ------------------------------------------------------------
_LAUNCH = re.compile(r"(\w+(?:<[^<>;]*>)?)<<<(.*), (\d+), 0, st>>>\((.*)\);")


class _HostHbm:
    """Stands in for the package in tests/test_hbm_kernels_gpu.py: debug_hbm() runs the kernels' host build."""
    WHICH = {"preproc": 0, "postproc": 1, "flow_tta_avg": 2, "warp": 3, "temporal_merge_v2": 4, "temporal_merge_v1": 5}

    def __init__(self, lib):
        self.lib = lib
        lib.emu_debug_hbm.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 3
        lib.emu_debug_hbm.restype = ctypes.c_int

    def debug_hbm(self, which, w, h, c, a, b, out):
        a = np.ascontiguousarray(a)
        b = None if b is None else np.ascontiguousarray(b)
        assert out.flags["C_CONTIGUOUS"]
        r = self.lib.emu_debug_hbm(self.WHICH[which], w, h, c, a.ctypes.data, None if b is None else b.ctypes.data, out.ctypes.data)
        assert r == 0


@pytest.fixture(scope="module")
def host_hbm(tmp_path_factory):
    inc = _cuda_include()
    if shutil.which("g++") is None or inc is None:
        pytest.skip("g++ or CUDA headers not available")
    d = tmp_path_factory.mktemp("emu_hbm")
    csrc = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "csrc")
    src = open(os.path.join(csrc, "hbm_kernels.cu")).read()
    out, n = _LAUNCH.subn(lambda m: "emu_launch(dim3(%s), %s, [&]() { %s(%s); });" % (m.group(2), m.group(3), m.group(1), m.group(4)), src)
    assert n == src.count("<<<") and n >= 15, (n, src.count("<<<"))  # every launch statement of the file was understood
    open(str(d / "hbm_kernels_emu.inc"), "w").write(out)
    so = str(d / "libemu_hbm.so")
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-I" + inc, "-I" + csrc, "-I" + str(d),
                        os.path.join(ROOT, "tests", "emu", "emu_hbm.cpp"), "-o", so], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]
    return _HostHbm(ctypes.CDLL(so))


def _gpu_cases(fn):
    """the (args) tuples of a parametrized test of tests/test_hbm_kernels_gpu.py"""
    for m in getattr(fn, "pytestmark", []):
        if m.name == "parametrize":
            return list(m.args[1])
    raise AssertionError("not parametrized")


def test_hbm_kernels_host_emulation(host_hbm):
    """Every case of tests/test_hbm_kernels_gpu.py (numpy restatements of the reference's CPU loops, partial tiles included), run
    against the kernels' host build: index arithmetic, tile transposes through `__shared__` + `__syncthreads()`, float4 paths and
    their scalar fall-backs are verified without a GPU.  -ffp-contract=off: no fused multiply-adds on the host, so the two
    last-bit allowances of the GPU tests are not even needed here (they still apply: the same assertions run)."""
    import test_hbm_kernels_gpu as g
    ran = 0
    for fn in (g.test_preproc, g.test_postproc, g.test_flow_tta_avg, g.test_warp, g.test_temporal_merge_v2, g.test_temporal_merge_v1):
        for case in _gpu_cases(fn):
            fn(host_hbm, *case)
            ran += 1
    assert ran >= 22
