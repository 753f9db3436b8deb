"""The HBM-side kernels of the generic path (csrc/hbm_kernels.cu: preproc, postproc, TTA flow average, temporal merges, warp)
one by one against numpy restatements of the reference's CPU loops (src/rife.cpp, src/warp.cpp; line refs per test), at
sizes that exercise partial 32 x 32 tiles.  The kernels do the reference's arithmetic in the reference's order: results are
bit-exact except where the compiler may contract a*b + c into one fused multiply-add (the final v*255 + 0.5 of postproc, the
two lerps of warp), which numpy rounds twice -- those two compare with a last-bit allowance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _orient(o, plane):
    """orientation o of a [hp][wp] plane (rife.cpp:3340-3364); 4-7 are [wp][hp]"""
    if o == 0:
        return plane
    if o == 1:
        return plane[:, ::-1]
    if o == 2:
        return plane[::-1, ::-1]
    if o == 3:
        return plane[::-1, :]
    t = plane.T
    if o == 4:
        return t
    if o == 5:
        return t[:, ::-1]
    if o == 6:
        return t[::-1, ::-1]
    return t[::-1, :]


def _unorient(o, plane):
    """inverse of _orient: back to [hp][wp]"""
    if o < 4:
        return _orient(o, plane)  # the flips are involutions
    if o == 4:
        return plane.T
    if o == 5:
        return plane[:, ::-1].T
    if o == 6:
        return plane[::-1, ::-1].T
    return plane[::-1, :].T


@pytest.mark.parametrize("w,h,norient", [(100, 70, 8), (96, 64, 1), (33, 31, 8), (257, 129, 8)])
def test_preproc(pkg, w, h, norient):
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    wp, hp = (w + 31) // 32 * 32, (h + 31) // 32 * 32
    out = np.full(norient * 3 * hp * wp, -1, np.float32)
    pkg.debug_hbm("preproc", w, h, norient, rgb, None, out)
    pad = np.zeros((3, hp, wp), np.float32)
    pad[:, :h, :w] = (rgb.astype(np.float32) * np.float32(1 / 255.0)).transpose(2, 0, 1)  # rife.cpp:4152-4211
    for o in range(norient):
        got = out[o * 3 * hp * wp:(o + 1) * 3 * hp * wp]
        want = np.stack([np.ascontiguousarray(_orient(o, pad[c])) for c in range(3)])
        assert np.array_equal(got, want.ravel()), o


@pytest.mark.parametrize("w,h,n_in", [(100, 70, 1), (96, 64, 2), (100, 70, 2), (100, 70, 8), (50, 90, 16), (128, 96, 16)])
def test_postproc(pkg, w, h, n_in):
    rng = np.random.default_rng(2)
    wp, hp = (w + 31) // 32 * 32, (h + 31) // 32 * 32
    planes = rng.uniform(-0.1, 1.1, (n_in, 3, hp, wp)).astype(np.float32)  # in orientation 0
    ins = np.empty((n_in, 3 * hp * wp), np.float32)
    for i in range(n_in):
        o = i & 7 if n_in >= 8 else 0
        ins[i] = np.stack([np.ascontiguousarray(_orient(o, planes[i, c])) for c in range(3)]).ravel()
    out = np.zeros((h, w, 3), np.uint8)
    pkg.debug_hbm("postproc", w, h, n_in, ins, None, out)
    f32 = np.float32
    if n_in == 1:
        v = planes[0] * f32(255) + f32(0.5)                                        # rife.cpp:4375-4398
    elif n_in == 2:
        v = (planes[0] + planes[1]) * f32(0.5) * f32(255) + f32(0.5)               # rife.cpp:4356-4371
    else:
        def mean8(p):                                                              # rife.cpp:4060-4144: added in orientation order, / 8
            s = np.zeros_like(p[0])
            for i in range(8):
                s = s + p[i]
            return s / f32(8)
        v = mean8(planes[:8]) * f32(255) + f32(0.5) if n_in == 8 else (mean8(planes[:8]) + mean8(planes[8:])) * f32(0.5) * f32(255) + f32(0.5)
    want = np.clip(v.astype(np.int32), 0, 255).astype(np.uint8)[:, :h, :w].transpose(1, 2, 0)  # (int) truncation, mat_pixel.cpp:158
    d = np.abs(out.astype(np.int16) - want.astype(np.int16))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3, (d.max(), (d > 0).mean())  # fma(v, 255, 0.5) vs two roundings, at truncation boundaries only


@pytest.mark.parametrize("fw,fh,nch", [(20, 12, 5), (52, 36, 4), (64, 32, 2), (33, 65, 5)])
def test_flow_tta_avg(pkg, fw, fh, nch):
    rng = np.random.default_rng(3)
    n = fw * fh
    blobs = rng.uniform(-4, 4, (8, nch, n)).astype(np.float32)
    out = np.empty_like(blobs)
    pkg.debug_hbm("flow_tta_avg", fw, fh, nch, blobs, None, out)
    # back to orientation 0 per blob and channel
    un = np.empty((8, nch, fh, fw), np.float32)
    for o in range(8):
        shp = (fh, fw) if o < 4 else (fw, fh)
        for c in range(nch):
            un[o, c] = _unorient(o, blobs[o, c].reshape(shp))
    f32 = np.float32
    want = np.empty_like(un)
    for k in range(2 if nch >= 4 else 1):  # rife.cpp:1541-1719 / 3515-3668
        cx, cy = 2 * k, 2 * k + 1
        x = (un[0, cx] + -un[1, cx] + -un[2, cx] + un[3, cx] + un[4, cy] + un[5, cy] + -un[6, cy] + -un[7, cy]) * f32(0.125)
        y = (un[0, cy] + un[1, cy] + -un[2, cy] + -un[3, cy] + un[4, cx] + -un[5, cx] + -un[6, cx] + un[7, cx]) * f32(0.125)
        sx = [1, -1, -1, 1]
        sy = [1, 1, -1, -1]
        for o in range(4):
            want[o, cx], want[o, cy] = sx[o] * x, sy[o] * y
        tx = [1, -1, -1, 1]
        ty = [1, 1, -1, -1]
        for o in range(4, 8):
            want[o, cx], want[o, cy] = tx[o - 4] * y, ty[o - 4] * x
    if nch == 5:
        m = np.zeros((fh, fw), np.float32)
        for o in range(8):
            m = m + un[o, 4]
        m = m * f32(0.125)
        for o in range(8):
            want[o, 4] = m
    for o in range(8):
        for c in range(nch):
            got = _unorient(o, out[o, c].reshape((fh, fw) if o < 4 else (fw, fh)))
            assert np.array_equal(got, want[o, c]), (o, c)


@pytest.mark.parametrize("w,h,c", [(100, 70, 3), (64, 48, 32), (37, 29, 13)])
def test_warp(pkg, w, h, c):
    rng = np.random.default_rng(4)
    img = rng.uniform(0, 1, (c, h, w)).astype(np.float32)
    flow = rng.uniform(-6, 6, (2, h, w)).astype(np.float32)
    flow[:, 0, 0] = [-1000, 1000]  # far outside: the clamps
    out = np.empty_like(img)
    pkg.debug_hbm("warp", w, h, c, img, flow, out)
    f32 = np.float32
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    sx, sy = xs + flow[0], ys + flow[1]                                            # warp.cpp:96-168
    x0 = np.floor(sx).astype(np.int64)
    y0 = np.floor(sy).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    x0, x1 = np.clip(x0, 0, w - 1), np.clip(x1, 0, w - 1)
    y0, y1 = np.clip(y0, 0, h - 1), np.clip(y1, 0, h - 1)
    al, be = sx - x0.astype(np.float32), sy - y0.astype(np.float32)               # alpha / beta AFTER clamping
    want = np.empty_like(img)
    for q in range(c):
        p = img[q]
        v4 = p[y0, x0] * (f32(1) - al) + p[y0, x1] * al
        v5 = p[y1, x0] * (f32(1) - al) + p[y1, x1] * al
        want[q] = v4 * (f32(1) - be) + v5 * be
    # fused multiply-adds in the two lerps: a last-bit allowance, scaled where a clamped tap makes |alpha| or |beta| large
    # (the lerp then extrapolates and cancellation magnifies the rounding difference)
    tol = np.float32(3e-7) * (1 + np.abs(al)) * (1 + np.abs(be)) * 4
    assert (np.abs(out - want) <= tol[None]).all(), np.abs(out - want).max()


@pytest.mark.parametrize("w,h,mask", [(20, 12, 1), (21, 11, 1), (64, 32, 0)])
def test_temporal_merge_v2(pkg, w, h, mask):
    rng = np.random.default_rng(5)
    n, nc = w * h, 4 + mask
    f = rng.uniform(-3, 3, (nc, n)).astype(np.float32)
    fr = rng.uniform(-3, 3, (nc, n)).astype(np.float32)
    out = np.empty((2, nc, n), np.float32)
    pkg.debug_hbm("temporal_merge_v2", w, h, mask, f, fr, out)
    h5 = np.float32(0.5)
    x, y, z, ww = (f[0] + fr[2]) * h5, (f[1] + fr[3]) * h5, (f[2] + fr[0]) * h5, (f[3] + fr[1]) * h5  # rife.cpp:2285-2306
    wf, wr = [x, y, z, ww], [z, ww, x, y]
    if mask:
        m = (f[4] - fr[4]) * h5                                                    # rife.cpp:4290-4311
        wf.append(m)
        wr.append(-m)
    assert np.array_equal(out[0], np.stack(wf)) and np.array_equal(out[1], np.stack(wr))


@pytest.mark.parametrize("w,h", [(20, 12), (21, 11)])
def test_temporal_merge_v1(pkg, w, h):
    rng = np.random.default_rng(6)
    n = w * h
    f = rng.uniform(-3, 3, (2, n)).astype(np.float32)
    fr = rng.uniform(-3, 3, (2, n)).astype(np.float32)
    out = np.empty((2, 2, n), np.float32)
    pkg.debug_hbm("temporal_merge_v1", w, h, 1, f, fr, out)
    x = (f - fr) * np.float32(0.5)                                                 # rife.cpp:2307-2319
    assert np.array_equal(out[0], x) and np.array_equal(out[1], -x)
