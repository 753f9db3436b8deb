// emu_fused.cpp -- TEST INFRASTRUCTURE.  Compiles the device code of the fused rife-v4.6 path
// (rife-ncnn-vulkan_b200/csrc/fused_v46_kernels.cuh) for the HOST: a launch becomes a loop over the grid, blockIdx /
// threadIdx are thread-local globals.  No GPU, no CUDA runtime.  It checks what can be checked without the conv stacks:
//   * the three storage variants of the full-resolution flow / mask planes (recompute_fm 0 / 1 / 2) feed the block heads
//     and the tail the same values: head tensors x2, x3 and the output frame must be bit-identical;
//   * the head / tail kernels against a direct restatement of SURVEY.md Appendix B written with whole-image loops
//     (upsample -> axpy -> warp -> downsample), i.e. the fusion (footprints, tap bookkeeping) changes nothing.
// Build: g++ -O1 -ffp-contract=off -I/usr/local/cuda/include -I<csrc> emu_fused.cpp   (tests/test_emu_cpu.py)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <cuda_fp16.h>
#include <vector_functions.h>
#include <vector_types.h>

// ---- the bits of the CUDA programming model the kernels use ----
static thread_local uint3 blockIdx, threadIdx;
static thread_local dim3 blockDim, gridDim;
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
#ifndef __grid_constant__
#define __grid_constant__
#endif

#define RIFE_FUSED_EMU 1
#include "fused_v46_kernels.cuh"

using namespace rife::fusedk;
using rife::V46_MAX_BATCH;

template <class K, class... A>
static void launch(dim3 g, dim3 b, K k, A... a) {
    gridDim = g;
    blockDim = b;
    for (unsigned z = 0; z < g.z; z++)
        for (unsigned y = 0; y < g.y; y++)
            for (unsigned x = 0; x < g.x; x++)
                for (unsigned t = 0; t < b.x; t++) {
                    blockIdx = make_uint3(x, y, z);
                    threadIdx = make_uint3(t, 0, 0);
                    k(a...);
                }
}
static unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }
static float frand(float lo, float hi) { return lo + (hi - lo) * (float)(rnd() & 0xffff) / 65535.f; }

// ---- whole-image restatement of the dataflow between the conv stacks ----
// ncnn Interp bilinear (interp.cpp:54-175): coefficients in double -> float, horizontal pass then vertical
static void ref_lin(int d, double scale, int in_n, int& s, float& f) {
    float fx = (float)((d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= in_n - 1) { sx = in_n - 2; fx = 1.f; }
    s = sx; f = fx;
}
static void ref_resize(const float* in, int ih, int iw, float* out, int oh, int ow) {
    if (ih == oh && iw == ow) { memcpy(out, in, sizeof(float) * ih * iw); return; }
    const double sy_ = (double)ih / oh, sx_ = (double)iw / ow;
    for (int y = 0; y < oh; y++) {
        int sy; float fy;
        ref_lin(y, sy_, ih, sy, fy);
        for (int x = 0; x < ow; x++) {
            int sx; float fx;
            ref_lin(x, sx_, iw, sx, fx);
            const float* r0 = in + (size_t)sy * iw + sx;
            const float* r1 = r0 + iw;
            float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
            float row0 = r0[0] * a0 + r0[1] * a1;
            float row1 = r1[0] * a0 + r1[1] * a1;
            out[(size_t)y * ow + x] = row0 * b0 + row1 * b1;
        }
    }
}
// src/warp.cpp:96-168 on one plane
static void ref_warp(const float* img, const float* fx, const float* fy, int h, int w, float* out) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float sx = x + fx[(size_t)y * w + x], sy = y + fy[(size_t)y * w + x];
            int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
            int x1 = x0 + 1, y1 = y0 + 1;
            x0 = std::min(std::max(x0, 0), w - 1); y0 = std::min(std::max(y0, 0), h - 1);
            x1 = std::min(std::max(x1, 0), w - 1); y1 = std::min(std::max(y1, 0), h - 1);
            float a = sx - x0, b = sy - y0;
            float v4 = img[(size_t)y0 * w + x0] * (1 - a) + img[(size_t)y0 * w + x1] * a;
            float v5 = img[(size_t)y1 * w + x0] * (1 - a) + img[(size_t)y1 * w + x1] * a;
            out[(size_t)y * w + x] = v4 * (1 - b) + v5 * b;
        }
}

struct Case {
    int w, h, wp, hp, n;
    bool v4 = false;                     // rife-v4 layout: 5 planes per block output at half the block resolution
    int packed = 0;                      // block-head tensors in the packed form (one plane, lo parts of the flow channels in slots 12..15)
    int contig = 0;                      // tail: 0 = crop the padded rows, 1 = the reference CPU path's contiguous read
    int dch() const { return v4 ? 5 : 6; }
    int up() const { return v4 ? 2 : 1; }
    std::vector<uchar4> rgbx;            // 2 * n frames, padded
    std::vector<float> d[4];             // per block: n x 6 planes at 1/8, 1/4, 1/2, 1/1
    TBatch tb;
    InBatch ib;
};

static void make_case(Case& c, int w, int h, int n, bool v4 = false, int contig = 0, int packed = 0) {
    c.w = w; c.h = h; c.n = n; c.v4 = v4; c.contig = contig; c.packed = packed;
    c.wp = (w + 31) / 32 * 32; c.hp = (h + 31) / 32 * 32;
    const size_t plane = (size_t)c.wp * c.hp;
    c.rgbx.assign(2 * n * plane, make_uchar4(0, 0, 0, 0));
    for (int f = 0; f < 2 * n; f++)
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                int base = (int)(120 + 60 * sinf(0.13f * (x + 2 * f)) * cosf(0.09f * (y - f)));
                c.rgbx[f * plane + (size_t)y * c.wp + x] = make_uchar4((unsigned char)std::min(255, base + (int)(rnd() & 31)), (unsigned char)std::min(255, base + (int)(rnd() & 15)),
                                                                        (unsigned char)std::max(0, base - (int)(rnd() & 31)), 0);
            }
    static const int S[4] = {8, 4, 2, 1};
    for (int k = 0; k < 4; k++) {
        const size_t hk = c.hp / S[k] / c.up(), wk = c.wp / S[k] / c.up();
        c.d[k].resize((size_t)n * c.dch() * hk * wk);
        // flow increments of a few pixels at this block's own scale, mask increments of order 1
        for (size_t i = 0; i < c.d[k].size(); i++) {
            const int ch = (int)((i / (hk * wk)) % c.dch());
            c.d[k][i] = ch < 4 ? frand(-1.5f, 1.5f) : frand(-2.f, 2.f);
        }
    }
    for (int b = 0; b < V46_MAX_BATCH; b++) {
        c.tb.t[b] = b < n ? 0.125f + 0.1f * b : 0.f;
        c.ib.p0[b] = b < n ? c.rgbx.data() + (size_t)(2 * b) * plane : nullptr;
        c.ib.p1[b] = b < n ? c.rgbx.data() + (size_t)(2 * b + 1) * plane : nullptr;
    }
}

struct Result {
    std::vector<__half> x[4];
    std::vector<uint8_t> out;
};

// the launches of V46Runner::run_batch that are not convolutions, with `rc` = the recompute_fm option
static void run_fused(Case& c, int rc, Result& r) {
    const int wp = c.wp, hp = c.hp, n = c.n;
    const size_t plane = (size_t)wp * hp;
    std::vector<float> F((size_t)n * 4 * plane, 1e30f), M((size_t)n * plane, 1e30f);  // poison: reading an unwritten plane shows
    static const int S[4] = {8, 4, 2, 1};
    for (int k = 0; k < 4; k++) r.x[k].assign((size_t)n * 16 * (hp / S[k]) * (wp / S[k]) * 2, __float2half_rn(-77.f));
    const bool rc2 = rc >= 2, rc1 = rc == 1;
    for (int k = 0; k < 4; k++) {
        const int hk = hp / S[k], wk = wp / S[k];
        dim3 g(cdiv(wk, 128), hk, n), b(128, 1, 1);
        float* d0 = c.d[0].data(); float* d1 = c.d[1].data(); float* d2 = c.d[2].data();
        __half* x = r.x[k].data();
        const int dch = c.dch();
        if (k == 0) launch(g, b, head0_kernel, c.ib, c.tb, hp, wp, hk, wk, x, c.packed);
        else if (c.v4) {
            if (k == 1) launch(g, b, head_update_kernel<4, 16, 0, 16, false>, c.ib, F.data(), M.data(), (const float*)d0, hp / 16, wp / 16, (const float*)nullptr, 0, 0, (const float*)nullptr, 0, 0, c.tb, hp, wp, hk, wk, x, dch, c.packed);
            else if (k == 2) launch(g, b, head_update_kernel<2, 8, 1, 16, true>, c.ib, F.data(), M.data(), (const float*)d1, hp / 8, wp / 8, (const float*)d0, hp / 16, wp / 16, (const float*)nullptr, 0, 0, c.tb, hp, wp, hk, wk, x, dch, c.packed);
            else launch(g, b, head_update_kernel<1, 4, 2, 16, true>, c.ib, F.data(), M.data(), (const float*)d2, hp / 4, wp / 4, (const float*)nullptr, 0, 0, (const float*)nullptr, 0, 0, c.tb, hp, wp, hk, wk, x, dch, c.packed);
        } else if (k == 1) launch(g, b, head_update_kernel<4, 8, 0, 8, false>, c.ib, F.data(), M.data(), (const float*)d0, hp / 8, wp / 8, (const float*)nullptr, 0, 0, (const float*)nullptr, 0, 0, c.tb, hp, wp, hk, wk, x, dch, c.packed);
        else if (k == 2) {
            if (rc2) launch(g, b, head_update_kernel<2, 4, 1, 8, false>, c.ib, F.data(), M.data(), (const float*)d1, hp / 4, wp / 4, (const float*)d0, hp / 8, wp / 8, (const float*)nullptr, 0, 0, c.tb, hp, wp, hk, wk, x, dch, c.packed);
            else launch(g, b, head_update_kernel<2, 4, 1, 8, true>, c.ib, F.data(), M.data(), (const float*)d1, hp / 4, wp / 4, (const float*)d0, hp / 8, wp / 8, (const float*)nullptr, 0, 0, c.tb, hp, wp, hk, wk, x, dch, c.packed);
        } else {
            if (rc2) launch(g, b, head_update_kernel<1, 2, 3, 4, false>, c.ib, F.data(), M.data(), (const float*)d2, hp / 2, wp / 2, (const float*)d1, hp / 4, wp / 4, (const float*)d0, hp / 8, wp / 8, c.tb, hp, wp, hk, wk, x, dch, c.packed);
            else if (rc1) launch(g, b, head_update_kernel<1, 2, 2, 8, false>, c.ib, F.data(), M.data(), (const float*)d2, hp / 2, wp / 2, (const float*)nullptr, 0, 0, (const float*)nullptr, 0, 0, c.tb, hp, wp, hk, wk, x, dch, c.packed);
            else launch(g, b, head_update_kernel<1, 2, 2, 8, true>, c.ib, F.data(), M.data(), (const float*)d2, hp / 2, wp / 2, (const float*)nullptr, 0, 0, (const float*)nullptr, 0, 0, c.tb, hp, wp, hk, wk, x, dch, c.packed);
        }
    }
    r.out.assign((size_t)n * c.w * c.h * 3, 0);
    OutBatch ob;
    for (int b = 0; b < V46_MAX_BATCH; b++) ob.p[b] = b < n ? r.out.data() + (size_t)b * c.w * c.h * 3 : nullptr;
    DSrc ds;
    ds.d[0] = c.d[0].data(); ds.d[1] = c.d[1].data(); ds.d[2] = c.d[2].data();
    dim3 tg(cdiv(c.w, 128), c.h, n), tbk(128, 1, 1);
    const float* Fc = F.data(); const float* Mc = M.data(); const float* d3 = c.d[3].data();
    if (c.v4) launch(tg, tbk, tail_kernel<0, 2>, c.ib, Fc, Mc, d3, hp, wp, ob, c.w, c.h, ds, c.contig, c.dch(), 0);
    else if (rc2) launch(tg, tbk, tail_kernel<2, 1>, c.ib, Fc, Mc, d3, hp, wp, ob, c.w, c.h, ds, c.contig, c.dch(), 0);
    else if (rc1) launch(tg, tbk, tail_kernel<1, 1>, c.ib, Fc, Mc, d3, hp, wp, ob, c.w, c.h, ds, c.contig, c.dch(), 0);
    else launch(tg, tbk, tail_kernel<0, 1>, c.ib, Fc, Mc, d3, hp, wp, ob, c.w, c.h, ds, c.contig, c.dch(), 0);
}

// value of channel ch of head tensor k (C8 space-to-depth, hi + lo planes) at output pixel (oy, ox)
static float head_at(const Result& r, int k, int b, int oh, int ow, int ch, int oy, int ox, int packed = 0) {
    const __half* base = r.x[k].data() + (size_t)b * 16 * oh * ow * 2;
    const size_t sub = (size_t)(oh >> 1) * (ow >> 1), pl = (size_t)16 * oh * ow;
    const int par = (oy & 1) * 2 + (ox & 1), g = ch >> 3, j = ch & 7;
    const size_t off = (((size_t)par * 2 + g) * sub + (size_t)(oy >> 1) * (ow >> 1) + (ox >> 1)) * 8 + j;
    if (packed) {  // one plane; slots 12..15 = lo parts of channels 8..11 (which the consuming conv weights like 8..11)
        if (ch >= 12) return 0.f;
        return __half2float(base[off]) + (ch >= 8 ? __half2float(base[off + 4]) : 0.f);
    }
    return __half2float(base[off]) + __half2float(base[pl + off]);
}

// whole-image reference of one pair: returns the head tensors (12 or 7 channels, fp32) and the output frame
static void run_reference(const Case& c, int b, std::vector<float> xr[4], std::vector<uint8_t>& out) {
    const int wp = c.wp, hp = c.hp;
    const size_t plane = (size_t)wp * hp;
    std::vector<float> I[2][3];
    for (int f = 0; f < 2; f++)
        for (int ch = 0; ch < 3; ch++) {
            I[f][ch].resize(plane);
            const uchar4* p = f == 0 ? c.ib.p0[b] : c.ib.p1[b];
            for (size_t i = 0; i < plane; i++) {
                const uchar4 q = p[i];
                I[f][ch][i] = (float)(ch == 0 ? q.x : (ch == 1 ? q.y : q.z)) * (1 / 255.f);
            }
        }
    std::vector<float> T(plane, c.tb.t[b]);
    static const int S[4] = {8, 4, 2, 1};
    std::vector<float> F[4], M(plane, 0.f);
    for (auto& f : F) f.assign(plane, 0.f);
    std::vector<float> tmp(plane), W[2][3];
    for (int k = 0; k < 4; k++) {
        const int hk = hp / S[k], wk = wp / S[k];
        const size_t pk = (size_t)hk * wk;
        if (k == 0) {
            xr[0].assign(7 * pk, 0.f);
            for (int ch = 0; ch < 3; ch++) ref_resize(I[0][ch].data(), hp, wp, &xr[0][ch * pk], hk, wk);
            for (int ch = 0; ch < 3; ch++) ref_resize(I[1][ch].data(), hp, wp, &xr[0][(3 + ch) * pk], hk, wk);
            ref_resize(T.data(), hp, wp, &xr[0][6 * pk], hk, wk);
        } else {
            // update after block k-1
            const int sp = S[k - 1] * c.up(), dh = hp / sp, dw = wp / sp;  // v4: 5 planes at half the block resolution, factor 2*S
            const float* d = c.d[k - 1].data() + (size_t)b * c.dch() * dh * dw;
            for (int ch = 0; ch < 5; ch++) {
                ref_resize(d + (size_t)ch * dh * dw, dh, dw, tmp.data(), hp, wp);
                for (size_t i = 0; i < plane; i++) {
                    if (ch < 4) F[ch][i] = k == 1 ? tmp[i] * (float)sp : F[ch][i] * 1.f + tmp[i] * (float)sp;
                    else M[i] = k == 1 ? tmp[i] : M[i] + tmp[i];
                }
            }
            for (int f = 0; f < 2; f++)
                for (int ch = 0; ch < 3; ch++) {
                    W[f][ch].resize(plane);
                    ref_warp(I[f][ch].data(), F[2 * f].data(), F[2 * f + 1].data(), hp, wp, W[f][ch].data());
                }
            xr[k].assign(12 * pk, 0.f);
            for (int f = 0; f < 2; f++)
                for (int ch = 0; ch < 3; ch++) ref_resize(W[f][ch].data(), hp, wp, &xr[k][(3 * f + ch) * pk], hk, wk);
            ref_resize(T.data(), hp, wp, &xr[k][6 * pk], hk, wk);
            ref_resize(M.data(), hp, wp, &xr[k][7 * pk], hk, wk);
            for (int ch = 0; ch < 4; ch++) {
                ref_resize(F[ch].data(), hp, wp, &xr[k][(8 + ch) * pk], hk, wk);
                if (S[k] > 1) for (size_t i = 0; i < pk; i++) xr[k][(8 + ch) * pk + i] = xr[k][(8 + ch) * pk + i] / (float)S[k];
            }
        }
    }
    // final update + blend + quantise (flownet.param:202-217, rife.cpp:4375-4398)
    if (!c.v4) {
        const float* d3 = c.d[3].data() + (size_t)b * 6 * plane;
        for (int ch = 0; ch < 4; ch++) for (size_t i = 0; i < plane; i++) F[ch][i] = F[ch][i] + d3[ch * plane + i];
        for (size_t i = 0; i < plane; i++) M[i] = M[i] + d3[4 * plane + i];
    } else {  // rife-v4: U = bilinear(flow3, 2); F = F*1 + U*2 (Eltwise), M = M + U[4]  (models/rife-v4/flownet.param:152-160)
        const int dh = hp / 2, dw = wp / 2;
        const float* d3 = c.d[3].data() + (size_t)b * 5 * dh * dw;
        for (int ch = 0; ch < 5; ch++) {
            ref_resize(d3 + (size_t)ch * dh * dw, dh, dw, tmp.data(), hp, wp);
            for (size_t i = 0; i < plane; i++) {
                if (ch < 4) F[ch][i] = F[ch][i] * 1.f + tmp[i] * 2.f;
                else M[i] = M[i] + tmp[i];
            }
        }
    }
    std::vector<float> O[3];
    for (int ch = 0; ch < 3; ch++) {
        std::vector<float> w0(plane), w1(plane);
        ref_warp(I[0][ch].data(), F[0].data(), F[1].data(), hp, wp, w0.data());
        ref_warp(I[1][ch].data(), F[2].data(), F[3].data(), hp, wp, w1.data());
        O[ch].resize(plane);
        for (size_t i = 0; i < plane; i++) {
            float m = M[i];
            m = fminf(m, 88.3762626647949f); m = fmaxf(m, -88.3762626647949f);
            m = 1.f / (1.f + expf(-m));
            O[ch][i] = w0[i] * m + w1[i] * (1.f - m);
        }
    }
    out.resize((size_t)c.w * c.h * 3);
    for (int y = 0; y < c.h; y++)
        for (int x = 0; x < c.w; x++) {
            // crop of the padded planes (rife_postproc.comp:42), or the reference CPU path's contiguous read (rife.cpp:4375-4387)
            const size_t i = (size_t)y * c.w + x, src = c.contig ? i : (size_t)y * wp + x;
            for (int ch = 0; ch < 3; ch++) {
                int iv = (int)(O[ch][src] * 255.f + 0.5f);
                out[i * 3 + ch] = (uint8_t)std::min(std::max(iv, 0), 255);
            }
        }
}

static uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv) {
    int fails = 0;
    uint64_t sum = 1469598103934665603ull;
    // lin_coeff of this build against the reference arithmetic (interp.cpp:54-91), exhaustively over the coordinates and scales in use
    {
        const double scales[6] = {8.0, 4.0, 2.0, 0.5, 0.25, 0.125};
        const int sizes[5] = {2, 3, 34, 1088, 8192};
        long bad = 0;
        for (double sc : scales)
            for (int in_n : sizes)
                for (int d = 0; d < 8192; d++) {
                    int s0, s1;
                    float f0, f1;
                    ref_lin(d, sc, in_n, s0, f0);
                    lin_coeff(d, sc, in_n, s1, f1);
                    if (s0 != s1 || memcmp(&f0, &f1, 4)) bad++;
                }
        printf("lin_coeff (RIFE_FUSED_LEAN=%d) vs reference arithmetic: %ld mismatches\n", RIFE_FUSED_LEAN, bad);
        if (bad) fails++;
    }
        // {w, h, pairs, rife-v4 layout, contiguous-read quirk, packed head tensors}
    const int sizes[][6] = {{64, 64, 2, 0, 0, 0}, {100, 70, 3, 0, 0, 0}, {100, 70, 2, 0, 1, 0}, {160, 96, 1, 0, 0, 0}, {96, 128, 2, 0, 0, 0}, {64, 64, 2, 1, 0, 0}, {100, 70, 2, 1, 0, 0},
                            {100, 70, 1, 1, 1, 0}, {100, 70, 2, 0, 0, 1}, {64, 64, 1, 1, 0, 1}};
    for (auto& sz : sizes) {
        Case c;
        make_case(c, sz[0], sz[1], sz[2], sz[3] != 0, sz[4], sz[5]);
        Result r0, r1, r2;
        run_fused(c, 0, r0);
        run_fused(c, c.v4 ? 0 : 1, r1);  // the recompute variants exist for the v4.6 layout only
        run_fused(c, c.v4 ? 0 : 2, r2);
        static const int S[4] = {8, 4, 2, 1};
        for (int k = 0; k < 4; k++) {
            const size_t nb = r0.x[k].size() * sizeof(__half);
            if (memcmp(r0.x[k].data(), r1.x[k].data(), nb)) { printf("FAIL %dx%d: head %d differs, recompute 1\n", sz[0], sz[1], k); fails++; }
            if (memcmp(r0.x[k].data(), r2.x[k].data(), nb)) { printf("FAIL %dx%d: head %d differs, recompute 2\n", sz[0], sz[1], k); fails++; }
        }
        if (r0.out != r1.out) { printf("FAIL %dx%d: output differs, recompute 1\n", sz[0], sz[1]); fails++; }
        if (r0.out != r2.out) { printf("FAIL %dx%d: output differs, recompute 2\n", sz[0], sz[1]); fails++; }
        // against the whole-image restatement
        double worst = 0;
        size_t odiff = 0, omax = 0;
        for (int b = 0; b < c.n; b++) {
            std::vector<float> xr[4];
            std::vector<uint8_t> oref;
            run_reference(c, b, xr, oref);
            for (int k = 0; k < 4; k++) {
                const int hk = c.hp / S[k], wk = c.wp / S[k], nch = k == 0 ? 7 : 12;
                for (int ch = 0; ch < 16; ch++)
                    for (int y = 0; y < hk; y++)
                        for (int x = 0; x < wk; x++) {
                            const float got = head_at(r0, k, b, hk, wk, ch, y, x, c.packed);
                            const float want = ch < nch ? xr[k][((size_t)ch * hk + y) * wk + x] : 0.f;
                            // hi + lo fp16 carries ~22 bits: compare with a relative 2^-20 / absolute 1e-6 allowance;
                            // the plain fp16 channels of the packed form carry 11: scale their error so that one bound serves both
                            double e = fabs((double)got - want) / std::max(1.0, fabs((double)want));
                            if (c.packed && ch < 8) e *= 2e-6 / 5e-4;
                            if (e > worst) worst = e;
                        }
            }
            const uint8_t* o = r0.out.data() + (size_t)b * c.w * c.h * 3;
            for (size_t i = 0; i < oref.size(); i++) {
                size_t dd = (size_t)abs((int)o[i] - (int)oref[i]);
                odiff += dd != 0;
                omax = std::max(omax, dd);
            }
        }
        for (int k = 0; k < 4; k++) sum = fnv(sum, r0.x[k].data(), r0.x[k].size() * sizeof(__half));
        sum = fnv(sum, r0.out.data(), r0.out.size());
        printf("%dx%d n=%d%s%s: head tensors vs restatement max rel err %.3g, output bytes differing %zu (max %zu)\n", sz[0], sz[1], sz[2], c.v4 ? " v4" : "", c.contig ? " contig" : (c.packed ? " packed" : ""), worst, odiff, omax);
        if (worst > 2e-6) { printf("FAIL %dx%d: head tensor mismatch\n", sz[0], sz[1]); fails++; }
        if (omax > 0) { printf("FAIL %dx%d: output mismatch\n", sz[0], sz[1]); fails++; }
    }
    printf("checksum of all head tensors and frames: %016llx\n", (unsigned long long)sum);
    printf(fails ? "EMU FAILED (%d)\n" : "EMU OK\n", fails);
    return fails ? 1 : 0;
}
