"""The fp32 kernels of the generic executor (csrc/generic_kernels.cu: the reference's ncnn layer semantics)
one by one against numpy restatements of the ncnn layer definitions, WITHOUT a GPU: tests/emu/emu_generic.cpp compiles the
unmodified kernel source for the host (a block's threads are real threads; __syncthreads / __shfl_xor_sync are barriers), only
the `<<<...>>>` launch statements are rewritten here.  On the GPU the same kernels are covered end to end by the precision-0
parity tests against the oracle; this file pins each operator by itself, partial tiles and broadcasting shapes included.
Sums are accumulated in a different order than numpy's (and with fmaf in the conv / inner product): those compare with a
relative tolerance, everything else bit for bit."""
import ctypes
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LAUNCH = re.compile(r"(\w+(?:<[^<>;]*>)?)<<<(.*), (\d+), (\w+), st>>>\((.*)\);")
F = np.float32
P = ctypes.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    inc = None
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            inc = os.path.join(d, "include")
    if shutil.which("g++") is None or inc is None:
        pytest.skip("g++ or CUDA headers not available")
    d = tmp_path_factory.mktemp("emu_generic")
    csrc = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "csrc")
    src = open(os.path.join(csrc, "generic_kernels.cu")).read()
    out, n = _LAUNCH.subn(lambda m: "emu_launch(dim3(%s), %s, %s, [&]() { %s(%s); });" % (m.group(2), m.group(3), m.group(4), m.group(1), m.group(5)), src)
    assert n == src.count("<<<") and n >= 14, (n, src.count("<<<"))  # every launch statement of the file was understood
    assert out.count("extern __shared__ float smem[];") == 1
    out = out.replace("extern __shared__ float smem[];", "float* smem = reinterpret_cast<float*>(emu_dyn_smem);")
    open(str(d / "generic_kernels_emu.inc"), "w").write(out)
    so = str(d / "libemu_generic.so")
    r = subprocess.run(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread", "-w", "-I" + inc, "-I" + csrc, "-I" + str(d),
                        os.path.join(ROOT, "tests", "emu", "emu_generic.cpp"), "-o", so], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]
    lib = ctypes.CDLL(so)
    i, f, z = ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    lib.emu_conv.argtypes = [P, i, i, i, P, P, i, i, i, i, i, f, f, P, i, f, P, P, i, i]
    lib.emu_deconv.argtypes = [P, i, i, i, P, P, i, i, f, P]
    lib.emu_unary.argtypes = [P, P, z, i, f, f]
    lib.emu_prelu.argtypes = [P, P, i, P, i, z]
    lib.emu_binary.argtypes = [P, i, z, P, i, z, P, i, z, i]
    lib.emu_eltwise_sum2.argtypes = [P, P, f, f, P, z]
    lib.emu_interp.argtypes = [P, i, i, i, P, i, i]
    lib.emu_pixelshuffle.argtypes = [P, i, i, i, P, i]
    lib.emu_avgpool.argtypes = [P, P, i, z, i]
    lib.emu_innerproduct.argtypes = [P, P, P, P, i, i, i, f]
    lib.emu_fill.argtypes = [P, z, f]
    return lib


def _act(v, act, p0=0.0, p1=0.0):
    """fused_activation.h:22-75"""
    if act == 1:
        return np.maximum(v, F(0))
    if act == 2:
        return np.where(v > 0, v, v * F(p0))
    if act == 3:
        return np.minimum(np.maximum(v, F(p0)), F(p1))
    if act == 4:
        v = np.clip(v, F(-88.3762626647949), F(88.3762626647949))
        return (F(1) / (F(1) + np.exp(-v))).astype(F)
    return v


def _conv_ref(x, w, bias, s, pad):
    """convolution.cpp:133-204: weights [oc][ic][ky][kx], zero padding"""
    cout, cin, k, _ = w.shape
    _, h, ww = x.shape
    oh, ow = (h + 2 * pad - k) // s + 1, (ww + 2 * pad - k) // s + 1
    xp = np.zeros((cin, h + 2 * pad, ww + 2 * pad), np.float64)
    xp[:, pad:pad + h, pad:pad + ww] = x
    out = np.zeros((cout, oh, ow), np.float64)
    for ky in range(k):
        for kx in range(k):
            out += np.einsum("oi,ihw->ohw", w[:, :, ky, kx].astype(np.float64), xp[:, ky:ky + s * (oh - 1) + 1:s, kx:kx + s * (ow - 1) + 1:s])
    if bias is not None:
        out += bias[:, None, None]
    return out.astype(F)


@pytest.mark.parametrize("cin,cout,k,s,pad,h,w", [(16, 64, 3, 1, 1, 20, 37), (7, 40, 3, 2, 1, 33, 50), (6, 24, 5, 2, 2, 40, 36), (24, 48, 5, 1, 2, 18, 34),
                                                   (13, 5, 3, 1, 1, 17, 33), (32, 12, 1, 1, 0, 9, 40), (8, 16, 2, 1, 0, 10, 35)])
@pytest.mark.parametrize("variant", ["plain", "leaky", "res_prelu"])
def test_conv_direct(emu, cin, cout, k, s, pad, h, w, variant):
    rng = np.random.default_rng(cin * 100 + cout + k)
    x = rng.uniform(-1, 1, (cin, h, w)).astype(F)
    wt = (rng.uniform(-1, 1, (cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(F)
    bias = rng.uniform(-0.5, 0.5, cout).astype(F)
    oh, ow = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    out = np.full((cout, oh, ow), np.nan, F)
    want = _conv_ref(x, wt, bias, s, pad)
    act, p0, res, post_act, post_p0, slope = 0, 0.0, None, 0, 0.0, None
    if variant == "leaky":
        act, p0 = 2, 0.2
        want = _act(want, 2, 0.2)
    elif variant == "res_prelu":  # v = conv + bias; v += res; per-channel PReLU (the fused epilogue of the generic executor)
        res = rng.uniform(-1, 1, (cout, oh, ow)).astype(F)
        slope = rng.uniform(0.05, 0.5, cout).astype(F)
        post_act = 5
        v = want + res
        want = np.where(v < 0, v * slope[:, None, None], v).astype(F)
    emu.emu_conv(_ptr(x), cin, h, w, _ptr(wt), _ptr(bias), cout, k, s, pad, act, p0, 0.0, _ptr(res), post_act, post_p0, _ptr(slope), _ptr(out), oh, ow)
    assert np.allclose(out, want, rtol=2e-5, atol=2e-6), np.abs(out - want).max()


@pytest.mark.parametrize("cin,cout,h,w", [(16, 24, 9, 20), (5, 5, 17, 33), (32, 70, 6, 7)])
def test_deconv4x4_s2(emu, cin, cout, h, w):
    """deconvolution.cpp:68-141 (scatter form, no kernel flip), pad 1: out[oc][2i + ky - 1][2j + kx - 1] += in[ic][i][j] * w[oc][ic][ky][kx]"""
    rng = np.random.default_rng(cin + cout)
    x = rng.uniform(-1, 1, (cin, h, w)).astype(F)
    wt = (rng.uniform(-1, 1, (cout, cin, 4, 4)) / np.sqrt(cin * 4)).astype(F)
    bias = rng.uniform(-0.5, 0.5, cout).astype(F)
    full = np.zeros((cout, 2 * h + 2, 2 * w + 2), np.float64)
    for ky in range(4):
        for kx in range(4):
            full[:, ky:ky + 2 * h:2, kx:kx + 2 * w:2] += np.einsum("oi,ihw->ohw", wt[:, :, ky, kx].astype(np.float64), x.astype(np.float64))
    want = (full[:, 1:-1, 1:-1] + bias[:, None, None]).astype(F)
    out = np.full((cout, 2 * h, 2 * w), np.nan, F)
    emu.emu_deconv(_ptr(x), cin, h, w, _ptr(wt), _ptr(bias), cout, 0, 0.0, _ptr(out))
    assert np.allclose(out, want, rtol=2e-5, atol=2e-6), np.abs(out - want).max()


def test_unary_ops(emu):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-4, 4, 1000), [0.0, -0.0, 100.0, -100.0, 1e-30]]).astype(F)
    p0, p1 = F(0.3), F(1.7)
    with np.errstate(divide="ignore"):
        rdiv = p0 / x
    cases = {0: np.maximum(x, F(0)), 1: np.where(x < 0, x * p0, x), 3: np.minimum(np.maximum(x, p0), p1), 4: -x, 5: x + p0, 6: x - p0, 7: x * p0, 8: x / p0,
             9: p0 - x, 10: rdiv, 11: x}
    for op, want in cases.items():  # kernels.h: U_RELU, U_LEAKY, U_SIGMOID, U_CLIP, U_NEG, U_ADD_S .. U_RDIV_S, U_COPY
        out = np.empty_like(x)
        emu.emu_unary(_ptr(x), _ptr(out), x.size, op, float(p0), float(p1))
        assert np.array_equal(out, want.astype(F), equal_nan=True), op
    out = np.empty_like(x)
    emu.emu_unary(_ptr(x), _ptr(out), x.size, 2, 0.0, 0.0)  # sigmoid.cpp:42-44 (clamped exponent); expf vs numpy's exp: last bits
    assert np.allclose(out, _act(x, 4), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("c,hw,nslope", [(5, 77, 5), (3, 128, 1)])
def test_prelu(emu, c, hw, nslope):
    rng = np.random.default_rng(2)
    x = rng.uniform(-2, 2, (c, hw)).astype(F)
    s = rng.uniform(0.05, 0.5, nslope).astype(F)
    out = np.empty_like(x)
    emu.emu_prelu(_ptr(x), _ptr(s), nslope, _ptr(out), c, hw)
    sl = s[:, None] if nslope > 1 else s[0]
    assert np.array_equal(out, np.where(x < 0, x * sl, x).astype(F))


@pytest.mark.parametrize("ac,ahw,bc,bhw", [("c", "hw", "c", "hw"), ("c", "hw", "c", 1), ("c", "hw", 1, "hw"), ("c", "hw", 1, 1), (1, "hw", "c", "hw"), ("c", 1, "c", "hw")])
@pytest.mark.parametrize("c,hw", [(6, 100), (5, 37), (3, 1)])
def test_binary_broadcasting(emu, c, hw, ac, ahw, bc, bhw):
    """binaryop.cpp:60-330, the shapes the models use: full, per-channel scalar, single plane, single value (either side)"""
    rng = np.random.default_rng(3)
    ac, bc = (c if ac == "c" else 1), (c if bc == "c" else 1)
    ahw, bhw = (hw if ahw == "hw" else 1), (hw if bhw == "hw" else 1)
    a = rng.uniform(0.5, 2, (ac, ahw)).astype(F)
    b = rng.uniform(0.5, 2, (bc, bhw)).astype(F)
    ops = {0: np.add, 1: np.subtract, 2: np.multiply, 3: np.divide, 4: np.maximum, 5: np.minimum, 6: np.power, 7: lambda x, y: y - x, 8: lambda x, y: y / x}
    for op, fn in ops.items():  # kernels.h: B_ADD, B_SUB, B_MUL, B_DIV, B_MAX, B_MIN, B_POW, B_RSUB, B_RDIV
        out = np.empty((c, hw), F)
        emu.emu_binary(_ptr(a), ac, ahw, _ptr(b), bc, bhw, _ptr(out), c, hw, op)
        want = np.broadcast_to(fn(a, b).astype(F), (c, hw))
        if op == 6:
            assert np.allclose(out, want, rtol=1e-6), op  # powf vs numpy
        else:
            assert np.array_equal(out, want), (op, ac, ahw, bc, bhw)


def test_eltwise_sum_with_coefficients_and_fill(emu):
    rng = np.random.default_rng(4)
    a, b = rng.uniform(-1, 1, 1003).astype(F), rng.uniform(-1, 1, 1003).astype(F)
    out = np.empty_like(a)
    emu.emu_eltwise_sum2(_ptr(a), _ptr(b), 0.25, -1.5, _ptr(out), a.size)  # eltwise.cpp:79-150
    assert np.array_equal(out, a * F(0.25) + b * F(-1.5))
    emu.emu_fill(_ptr(out), out.size, 0.375)
    assert np.array_equal(out, np.full_like(a, 0.375))


def _lin(d, in_n, out_n):
    """interp.cpp:54-91: coefficients in double, rounded to float"""
    fx = F((d + 0.5) * (float(in_n) / out_n) - 0.5)
    sx = int(np.floor(fx))
    fx = F(fx - F(sx))
    if sx < 0:
        sx, fx = 0, F(0)
    if sx >= in_n - 1:
        sx, fx = in_n - 2, F(1)
    return sx, fx


@pytest.mark.parametrize("c,h,w,oh,ow", [(3, 10, 14, 20, 28), (18, 16, 24, 8, 12), (2, 9, 13, 36, 52), (5, 12, 20, 3, 5)])
def test_interp_bilinear(emu, c, h, w, oh, ow):
    """interp.cpp:92-175: horizontal pass then vertical pass, float"""
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (c, h, w)).astype(F)
    out = np.empty((c, oh, ow), F)
    emu.emu_interp(_ptr(x), c, h, w, _ptr(out), oh, ow)
    want = np.empty_like(out)
    for oy in range(oh):
        sy, fy = _lin(oy, h, oh)
        for ox in range(ow):
            sx, fx = _lin(ox, w, ow)
            r0 = x[:, sy, sx] * (F(1) - fx) + x[:, sy, sx + 1] * fx
            r1 = x[:, sy + 1, sx] * (F(1) - fx) + x[:, sy + 1, sx + 1] * fx
            want[:, oy, ox] = r0 * (F(1) - fy) + r1 * fy
    assert np.array_equal(out, want)


@pytest.mark.parametrize("c,h,w,r", [(24, 7, 9, 2), (9, 5, 6, 3)])
def test_pixelshuffle(emu, c, h, w, r):
    """pixelshuffle.cpp:33-80 mode 0"""
    x = np.arange(c * h * w, dtype=F).reshape(c, h, w)
    out = np.empty((c // (r * r), h * r, w * r), F)
    emu.emu_pixelshuffle(_ptr(x), c, h, w, _ptr(out), r)
    want = x.reshape(c // (r * r), r, r, h, w).transpose(0, 3, 1, 4, 2).reshape(out.shape)
    assert np.array_equal(out, want)


@pytest.mark.parametrize("c,hw,scratch", [(7, 1000, 0), (3, 16384, 1), (4, 16390, 1), (2, 33, 1)])
def test_global_avgpool(emu, c, hw, scratch):
    """pooling.cpp:61-105; the two-stage form (16 slices per channel) starts at 16384 elements per plane"""
    rng = np.random.default_rng(6)
    x = rng.uniform(-1, 1, (c, hw)).astype(F)
    out = np.empty(c, F)
    emu.emu_avgpool(_ptr(x), _ptr(out), c, hw, scratch)
    assert np.allclose(out, x.astype(np.float64).mean(axis=1), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("nin,nout,act", [(48, 6, 1), (100, 13, 4), (7, 3, 0)])
def test_innerproduct(emu, nin, nout, act):
    """innerproduct.cpp: out[p] = act(bias[p] + sum_i w[p][i] x[i])"""
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, nin).astype(F)
    w = rng.uniform(-1, 1, (nout, nin)).astype(F)
    b = rng.uniform(-1, 1, nout).astype(F)
    out = np.empty(nout, F)
    emu.emu_innerproduct(_ptr(x), _ptr(w), _ptr(b), _ptr(out), nin, nout, act, 0.0)
    want = _act((w.astype(np.float64) @ x.astype(np.float64) + b).astype(F), act)
    assert np.allclose(out, want, rtol=1e-5, atol=1e-6)
