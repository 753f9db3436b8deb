"""CPU replay of the tcgen05 conv's operand views (csrc/tc_conv.cu): a numpy model of the no-swizzle K-major shared-memory
descriptors (row i of an operand = 16 contiguous bytes at start + i*16, second K half LBO bytes further) applied to
(a) the real packed weights (rife_b200_debug_pack_weights) and (b) a halo tile laid out as TMA delivers it, with the
exact start-address / TMEM-column arithmetic of the kernel's issue loops.  The unpaired issue order is the one verified
on the GPU (tests/test_tc_conv_gpu.py); the paired order (RIFE_B200_PAIR: one 2N-column MMA feeds the dy=2 tap of
accumulator j-1 and the dy=0 tap of accumulator j) must produce the same accumulators, and both must equal a direct
3x3 convolution / 4x4 stride-2 deconvolution of the tile."""
import ctypes

import numpy as np
import pytest

TWP = 64


def _pack(pkg, mode, w, cout, cin, N, ocs, paired):
    L = pkg.lib()
    n = (cin // 16) * 9 * 2 * N * 8
    out = np.zeros(n, np.uint16)
    w = np.ascontiguousarray(w, np.float32)
    r = L.rife_b200_debug_pack_weights(mode, cout, cin, N, ocs, paired, w.ctypes.data, out.ctypes.data, n)
    assert r == 0
    return out.view(np.float16).astype(np.float32)


def _a_view(slab, rows, start16):
    """A operand (128 x 16) from one K chunk's slab [2 halves][rows][64 px][8]: row i = pixel start16 + i of each half."""
    flat = slab.reshape(2, rows * TWP, 8)
    pad = np.zeros((2, 128 + 4 * TWP, 8), np.float32)  # overrun pad behind the slab (zeros here; those rows are halo columns)
    flat = np.concatenate([flat, pad], axis=1)
    return np.concatenate([flat[0, start16:start16 + 128], flat[1, start16:start16 + 128]], axis=1)  # [128][16]


def _b_view(wchunk, start16, lbo16, n):
    """B operand (n x 16): row r = 16 B at (start16 + r), second K half lbo16 rows further; wchunk = fp16 values of one K chunk"""
    rows = wchunk.reshape(-1, 8)
    return np.concatenate([rows[start16:start16 + n], rows[start16 + lbo16:start16 + lbo16 + n]], axis=1)  # [n][16]


def _accumulate(slabs, wpk, N, MT, paired):
    """Replays the mainloop of one tile: returns acc [MT][128][N]"""
    rows = 2 * MT + 2
    kcs = len(slabs)
    acc = np.zeros((128, MT * N), np.float64)
    wk = wpk.reshape(kcs, -1)
    rowstep = 2 * TWP
    for kc in range(kcs):
        slab = slabs[kc]
        if not paired:
            for tap in range(9):
                dy, dx = divmod(tap, 3)
                B = _b_view(wk[kc], tap * 2 * N, N, N)
                for m in range(MT):
                    A = _a_view(slab, rows, dy * TWP + dx + m * rowstep)
                    acc[:, m * N:(m + 1) * N] += A @ B.T
        else:
            for dx in range(3):
                blk = dx * 2 * 3 * N  # 16-byte rows from the chunk start
                for j in range(MT + 1):  # even views (umma_issue_pair)
                    A = _a_view(slab, rows, dx + j * rowstep)
                    if j == 0:
                        acc[:, 0:N] += A @ _b_view(wk[kc], blk + N, 3 * N, N).T
                    elif j == MT:
                        acc[:, (MT - 1) * N:MT * N] += A @ _b_view(wk[kc], blk, 3 * N, N).T
                    else:
                        acc[:, (j - 1) * N:(j + 1) * N] += A @ _b_view(wk[kc], blk, 3 * N, 2 * N).T
                for m in range(MT):  # odd views (umma_issue_tap at one row down, dy = 1 rows of the block)
                    A = _a_view(slab, rows, dx + TWP + m * rowstep)
                    acc[:, m * N:(m + 1) * N] += A @ _b_view(wk[kc], blk + 2 * N, 3 * N, N).T
    return acc.reshape(128, MT, N).transpose(1, 0, 2)


def _tile(cin, MT, rng):
    rows = 2 * MT + 2
    x = rng.standard_normal((cin, rows, TWP)).astype(np.float16).astype(np.float32)  # halo tile, planar
    slabs = [x[kc * 16:(kc + 1) * 16].reshape(2, 8, rows, TWP).transpose(0, 2, 3, 1).copy() for kc in range(cin // 16)]
    return x, slabs


@pytest.mark.parametrize("cin,cout,N,MT", [(64, 64, 64, 4), (32, 32, 32, 4), (96, 96, 96, 2), (128, 128, 128, 2), (48, 40, 48, 4)])
def test_conv3x3_views_paired_equals_unpaired_equals_direct(pkg, cin, cout, N, MT):
    rng = np.random.default_rng(cin * 7 + N)
    w = (rng.standard_normal((cout, cin, 3, 3)) * 0.1).astype(np.float16).astype(np.float32)
    x, slabs = _tile(cin, MT, rng)
    ref = np.zeros((MT, 128, N))
    for m in range(MT):  # position p of accumulator m: output row 2m + p // 64 (tile rows), column p % 64 (valid < 62)
        for p in range(128):
            r, c = 2 * m + p // TWP, p % TWP
            if c >= TWP - 2:
                continue
            patch = x[:, r:r + 3, c:c + 3]  # halo tile coordinates: output (r, c) reads rows r..r+2, cols c..c+2
            ref[m, p, :cout] = np.tensordot(w, patch, axes=([1, 2, 3], [0, 1, 2]))
    valid = (np.arange(128) % TWP) < TWP - 2
    for paired in (0, 1):
        acc = _accumulate(slabs, _pack(pkg, 0, w, cout, cin, N, 0, paired), N, MT, bool(paired))
        assert np.allclose(acc[:, valid], ref[:, valid], atol=1e-3), (paired, np.abs(acc[:, valid] - ref[:, valid]).max())


def test_deconv_views_paired_equals_unpaired(pkg):
    """deconv4x4 s2 as a 3x3-neighbourhood GEMM with N = 4 parities x ocs columns (flow head: 24 -> N = 96)"""
    cin, cout, ocs, N, MT = 64, 24, 24, 96, 2
    rng = np.random.default_rng(5)
    w = (rng.standard_normal((cout, cin, 4, 4)) * 0.1).astype(np.float16).astype(np.float32)
    x, slabs = _tile(cin, MT, rng)
    a0 = _accumulate(slabs, _pack(pkg, 1, w, cout, cin, N, ocs, 0), N, MT, False)
    a1 = _accumulate(slabs, _pack(pkg, 1, w, cout, cin, N, ocs, 1), N, MT, True)
    valid = (np.arange(128) % TWP) < TWP - 2
    assert np.allclose(a0[:, valid], a1[:, valid], atol=1e-3)
    # direct: out(2y+py, 2x+px, oc) = sum_{ic, ky, kx} in(iy, ix) w[oc][ic][ky][kx] with 2*iy - 1 + ky = 2y+py (ncnn deconvolution, pad 1)
    m, p = 1, 70
    r, c = 2 * m + p // TWP, p % TWP  # input pixel (y, x) = halo (r + 1, c + 1)
    for par in range(4):
        py, px = par >> 1, par & 1
        for oc in range(cout):
            s = 0.0
            for ky in range(4):
                for kx in range(4):
                    ty, tx = 2 * (r + 1) + py + 1 - ky, 2 * (c + 1) + px + 1 - kx
                    if ty % 2 or tx % 2:
                        continue
                    s += float(np.dot(x[:, ty // 2, tx // 2], w[oc, :, ky, kx]))
            assert abs(a1[m, p, par * ocs + oc] - s) < 1e-3, (par, oc, a1[m, p, par * ocs + oc], s)


# ---- wide tiles (kernel template WIDE): one-row accumulators of 128 pixels, [dy2 | dy1 | dy0] weight blocks ------------------
WTW = 128


def _a_view_wide(slab, rows, start16):
    """A operand (128 x 16) from a wide slab [2 halves][rows][128 px][8]: 128 consecutive pixels from start16 (wraps into the
    next row for dx > 0; behind the last row: the overrun pad)."""
    flat = slab.reshape(2, rows * WTW, 8)
    flat = np.concatenate([flat, np.zeros((2, 64, 8), np.float32)], axis=1)
    return np.concatenate([flat[0, start16:start16 + 128], flat[1, start16:start16 + 128]], axis=1)


def _accumulate_wide(slabs, wpk, N, MT):
    """umma_issue_wide (tools/gen_mma_issue.py wide_block): per kernel column dx, halo row r feeds accumulators r - dy."""
    rows = MT + 2
    kcs = len(slabs)
    acc = np.zeros((128, MT * N), np.float64)
    wk = wpk.reshape(kcs, -1)
    for kc in range(kcs):
        for dx in range(3):
            blk = dx * 2 * 3 * N
            for r in range(MT + 2):
                dy_max, dy_min = min(2, r), max(0, r - (MT - 1))
                ncol = (dy_max - dy_min + 1) * N
                A = _a_view_wide(slabs[kc], rows, dx + r * WTW)
                B = _b_view(wk[kc], blk + (2 - dy_max) * N, 3 * N, ncol)
                c0 = (r - dy_max) * N
                acc[:, c0:c0 + ncol] += A @ B.T
    return acc.reshape(128, MT, N).transpose(1, 0, 2)


@pytest.mark.parametrize("cin,cout,N,MT", [(64, 64, 64, 4), (32, 64, 64, 4)])
def test_conv3x3_wide_views_equal_direct(pkg, cin, cout, N, MT):
    rng = np.random.default_rng(cin + 3 * N)
    w = (rng.standard_normal((cout, cin, 3, 3)) * 0.1).astype(np.float16).astype(np.float32)
    rows = MT + 2
    x = rng.standard_normal((cin, rows, WTW)).astype(np.float16).astype(np.float32)
    slabs = [x[kc * 16:(kc + 1) * 16].reshape(2, 8, rows, WTW).transpose(0, 2, 3, 1).copy() for kc in range(cin // 16)]
    acc = _accumulate_wide(slabs, _pack(pkg, 0, w, cout, cin, N, 0, 2), N, MT)
    for m in range(MT):  # position p of accumulator m: output row m, column p (valid < 126)
        for p in range(0, WTW - 2, 7):
            patch = x[:, m:m + 3, p:p + 3]
            want = np.tensordot(w, patch, axes=([1, 2, 3], [0, 1, 2]))
            assert np.allclose(acc[m, p, :cout], want, atol=1e-3), (m, p)
