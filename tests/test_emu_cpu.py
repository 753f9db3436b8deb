"""Host emulation of the fused rife-v4.6 HBM kernels (tests/emu/emu_fused.cpp): the device code of
fused_v46_kernels.cuh compiled with g++ (grid = loops), checked (a) across the three flow / mask storage variants
(recompute_fm 0 / 1 / 2: bit-identical head tensors and frames) and (b) against a whole-image restatement of
SURVEY.md Appendix B.  Needs only g++ and the CUDA headers (cuda_fp16.h, vector_types.h); no GPU, no CUDA runtime."""
import os
import shutil
import subprocess

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
