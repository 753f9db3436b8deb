"""The layout conversions between the generic executor's planar fp32 blobs and the tensor-core path's tensors (csrc/tc_conv.cu:
planar_to_c8_kernel / c8_to_planar_kernel) on the HOST (tests/emu/emu_layout.cpp), against a numpy statement of the layout
(DESIGN.md section 3): C8 planar `[plane][Cpad/8][H][W][8]` fp16, plane 0 = hi = fp16(v), plane 1 = lo = fp16(v - hi); stride-2
consumers read the space-to-depth form `[py*2+px][Cpad/8][H/2][W/2][8]`; channels beyond C are zero.  Bit for bit."""
import ctypes
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LAUNCH = re.compile(r"(\w+)<<<(.*), (\d+), 0, st>>>\((.*)\);")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    inc = "/usr/local/cuda/include" if os.path.exists("/usr/local/cuda/include/cuda_fp16.h") else None
    if os.environ.get("CUDA_HOME") and os.path.exists(os.path.join(os.environ["CUDA_HOME"], "include", "cuda_fp16.h")):
        inc = os.path.join(os.environ["CUDA_HOME"], "include")
    if shutil.which("g++") is None or inc is None:
        pytest.skip("g++ or CUDA headers not available")
    d = tmp_path_factory.mktemp("emu_layout")
    csrc = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "csrc")
    src = open(os.path.join(csrc, "tc_conv.cu")).read()
    a, b = src.index("// ---- layout conversion kernels"), src.index("// ---- weight packing (host)")
    assert 0 < a < b
    pack2 = [l for l in src.splitlines() if l.startswith("__device__ __forceinline__ uint32_t pack2(")]
    assert len(pack2) == 1
    sec, n = _LAUNCH.subn(lambda m: "emu_launch(dim3(%s), %s, [&]() { %s(%s); });" % (m.group(2), m.group(3), m.group(1), m.group(4)), src[a:b])
    assert n == src[a:b].count("<<<") == 2
    open(str(d / "layout_kernels_emu.inc"), "w").write(pack2[0] + "\n" + sec)
    so = str(d / "libemu_layout.so")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-I" + inc, "-I" + str(d), os.path.join(ROOT, "tests", "emu", "emu_layout.cpp"), "-o", so],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]
    lib = ctypes.CDLL(so)
    lib.emu_planar_to_c8.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6
    lib.emu_c8_to_planar.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 6
    return lib


def _c8_ref(x, cpad, split, s2d):
    """[planes][...] fp16 bit patterns of the C8 tensor of planar fp32 x[C][H][W]"""
    c, h, w = x.shape
    xp = np.zeros((cpad, h, w), np.float32)
    xp[:c] = x
    hi = xp.astype(np.float16)
    lo = (xp - hi.astype(np.float32)).astype(np.float16)

    def lay(p):
        g = p.reshape(cpad // 8, 8, h, w).transpose(0, 2, 3, 1)  # [C/8][H][W][8]
        if s2d:  # [py*2+px][C/8][H/2][W/2][8]
            g = g.reshape(cpad // 8, h // 2, 2, w // 2, 2, 8).transpose(2, 4, 0, 1, 3, 5)
        return np.ascontiguousarray(g).ravel()
    return np.stack([lay(hi), lay(lo)]) if split else lay(hi)[None]


_SHAPES = [(12, 16, 6, 10), (7, 16, 4, 6), (64, 64, 5, 9), (8, 8, 2, 2), (3, 8, 8, 12)]


@pytest.mark.parametrize("c,cpad,h,w,s2d", [s + (0,) for s in _SHAPES] + [s + (1,) for s in _SHAPES if s[2] % 2 == 0 and s[3] % 2 == 0])  # space-to-depth: even sizes
@pytest.mark.parametrize("split", [1, 0])
def test_planar_to_c8_and_back(emu, c, cpad, h, w, s2d, split):
    rng = np.random.default_rng(c * 7 + h)
    x = (rng.standard_normal((c, h, w)) * rng.choice([1e-3, 1.0, 300.0], (c, 1, 1))).astype(np.float32)
    n = cpad * h * w
    out = np.full((2 if split else 1) * n, 0x7E00, np.uint16)  # NaN pattern: every element must be written
    emu.emu_planar_to_c8(x.ctypes.data, out.ctypes.data, c, h, w, split, cpad, s2d)
    want = _c8_ref(x, cpad, split, s2d).view(np.uint16).ravel()
    assert np.array_equal(out, want)
    back = np.full((c, h, w), np.nan, np.float32)
    emu.emu_c8_to_planar(out.ctypes.data, back.ctypes.data, c, h, w, split, cpad, s2d)
    hi = x.astype(np.float16).astype(np.float32)
    lo = (x - hi).astype(np.float16).astype(np.float32)
    assert np.array_equal(back, hi + lo if split else hi)
    if split:  # the split form carries ~22 bits: what makes tier 1 "fp32-equivalent" operands
        assert np.all(np.abs(back - x) <= np.abs(x) * 2.0 ** -20 + 1e-7)
