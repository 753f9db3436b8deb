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


def test_fused_kernels_host_emulation(tmp_path):
    inc = _cuda_include()
    if shutil.which("g++") is None or inc is None:
        pytest.skip("g++ or CUDA headers not available")
    exe = str(tmp_path / "emu_fused")
    src = os.path.join(ROOT, "tests", "emu", "emu_fused.cpp")
    csrc = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "csrc")
    # -ffp-contract=off: every variant must round identically; the comparison is bit-exact
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-I" + inc, "-I" + csrc, src, "-o", exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "EMU OK" in r.stdout, r.stdout[-4000:]
