// fused_v46_kernels.cuh -- device code of the fused rife-v4.6 path (see fused_v46.cu for the dataflow).  Kept in a header so
// that tests/emu_fused.cpp can compile the very same kernels for the host (thread loops instead of a grid) and check
// index arithmetic and the store / recompute variants against each other without a GPU.
#pragma once
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#ifndef RIFE_FUSED_EMU
#include "fused_v46.h"
#else
namespace rife { constexpr int V46_MAX_BATCH = 8; }  // host emulation (tests/emu): no runtime headers
#endif

namespace rife {
namespace fusedk {


// interp.cpp:54-91 (coefficient in double, rounded to float)
// `scale` = in_n / out_n, which on this path is always an exact power of two (8, 4, 2, 1/2, 1/4, 1/8) known at compile
// time, so the reference's double division is folded; the rest of the arithmetic is unchanged
// RIFE_FUSED_LEAN = 1 (compile-time, default 0 until measured on the GPU): same results bit for bit with fewer
// instructions on the slow pipes -- lin_coeff in integers instead of fp64 (every intermediate of the original is exactly
// representable, so nothing is rounded either way), packed float->half conversions in the hi/lo split.  tests/emu checks
// both builds against the whole-image restatement and against each other, exhaustively for lin_coeff.
#ifndef RIFE_FUSED_LEAN
#define RIFE_FUSED_LEAN 0
#endif
__device__ __forceinline__ void lin_coeff(int d, double scale, int in_n, int& s, float& f) {
#if RIFE_FUSED_LEAN
    int sx;
    float fx;
    if (scale >= 1.0) {  // down-sampling by k = 2, 4, 8: (d + 0.5) * k - 0.5 = d*k + (k/2 - 1) + 0.5
        const int k = (int)scale;
        sx = d * k + (k / 2 - 1);
        fx = 0.5f;
    } else {             // up-sampling by k: (d + 0.5) / k - 0.5 = (2d + 1 - k) / 2k
        const int k2 = 2 * (int)(1.0 / scale + 0.5);  // 2k, a power of two
        const int n = 2 * d + 1 - k2 / 2;
        sx = n >= 0 ? n / k2 : -((k2 - 1 - n) / k2);   // floor(n / 2k)
        fx = (float)(n - sx * k2) * (1.f / (float)k2);    // exact: numerator < 2k, 1/2k a power of two
    }
#else
    float fx = (float)((d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
#endif
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= in_n - 1) { sx = in_n - 2; fx = 1.f; }
    s = sx;
    f = fx;
}
__device__ __forceinline__ float bilerp(const float* __restrict__ p, int w, int sy, int sx, float a0, float a1, float b0, float b1) {
    const float* r0 = p + (size_t)sy * w + sx;
    const float* r1 = r0 + w;
    float row0 = r0[0] * a0 + r0[1] * a1;  // interp.cpp:92-175: horizontal pass, then vertical
    float row1 = r1[0] * a0 + r1[1] * a1;
    return row0 * b0 + row1 * b1;
}
// One frame, padded to wp x hp and widened to 4 bytes per pixel (R, G, B, 0; zeros in the pad region) by rgbx_kernel:
// a tap is one aligned 32-bit load.  The padded planar float image the reference works on (rife_preproc: v * 1/255)
// is evaluated on the fly.
struct Frame {
    const uchar4* p;
    int wp;
};
__device__ __forceinline__ void px3(const Frame& f, int x, int y, float* o) {
#if defined(__CUDA_ARCH__) && !defined(RIFE_FUSED_NO_PRMT)
    // u8 -> float without the quarter-rate conversion pipe (24 I2F.U8 per pixel made these kernels conversion-bound):
    // byte b dropped into the mantissa of 2^23 gives the float 2^23 + b exactly (one PRMT); then
    // fma(2^23 + b, r, -2^23 * r) = b * r rounded once -- bit-identical to (float)b * r, since 2^23 * r is exact.
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(f.p + (size_t)y * f.wp + x));
    constexpr float r = 1 / 255.f, off = -8388608.f * r;
    o[0] = __fmaf_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440)), r, off);
    o[1] = __fmaf_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441)), r, off);
    o[2] = __fmaf_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7442)), r, off);
#else
    const uchar4 q = __ldg(f.p + (size_t)y * f.wp + x);
    o[0] = (float)q.x * (1 / 255.f);
    o[1] = (float)q.y * (1 / 255.f);
    o[2] = (float)q.z * (1 / 255.f);
#endif
}
// bilinear tap of the three colour planes (interp.cpp:92-175: horizontal pass, then vertical)
__device__ __forceinline__ void bilerp3(const Frame& f, int sy, int sx, float a0, float a1, float b0, float b1, float* o) {
    float p00[3], p01[3], p10[3], p11[3];
    px3(f, sx, sy, p00); px3(f, sx + 1, sy, p01); px3(f, sx, sy + 1, p10); px3(f, sx + 1, sy + 1, p11);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float row0 = p00[c] * a0 + p01[c] * a1;
        float row1 = p10[c] * a0 + p11[c] * a1;
        o[c] = row0 * b0 + row1 * b1;
    }
}
// src/warp.cpp:96-168 for one pixel: taps and weights (alpha / beta taken after clamping)
struct WarpTap {
    int x0, x1, y0, y1;
    float a, b;
};
__device__ __forceinline__ WarpTap warp_tap(int x, int y, float fx, float fy, int w, int h) {
    float sx = x + fx, sy = y + fy;
    int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), w - 1);
    y0 = min(max(y0, 0), h - 1);
    x1 = min(max(x1, 0), w - 1);
    y1 = min(max(y1, 0), h - 1);
    WarpTap t;
    t.a = sx - x0;
    t.b = sy - y0;
    t.x0 = x0; t.x1 = x1; t.y0 = y0; t.y1 = y1;
    return t;
}
__device__ __forceinline__ void warp_sample3(const Frame& f, const WarpTap& t, float* o) {
    float p00[3], p01[3], p10[3], p11[3];
    px3(f, t.x0, t.y0, p00); px3(f, t.x1, t.y0, p01); px3(f, t.x0, t.y1, p10); px3(f, t.x1, t.y1, p11);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float v4 = p00[c] * (1 - t.a) + p01[c] * t.a;
        float v5 = p10[c] * (1 - t.a) + p11[c] * t.a;
        o[c] = v4 * (1 - t.b) + v5 * t.b;
    }
}
__device__ __forceinline__ uint32_t pack2h(__half a, __half b) { return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16); }

// writes 16 channel values of output pixel (oy, ox) (image oh x ow) as split fp16 into the space-to-depth C8 tensor
// [plane][py*2+px][2 groups][oh/2][ow/2][8]
__device__ __forceinline__ void store_c8_s2d_16(__half* out, const float* v, int oy, int ox, int oh, int ow) {
    const size_t sub = (size_t)(oh >> 1) * (ow >> 1);
    const size_t plane = (size_t)16 * oh * ow;
    const int par = (oy & 1) * 2 + (ox & 1);
    const size_t pix = (size_t)(oy >> 1) * (ow >> 1) + (ox >> 1);
#pragma unroll
    for (int g = 0; g < 2; g++) {
        const size_t off = (((size_t)par * 2 + g) * sub + pix) * 8;
#if RIFE_FUSED_LEAN
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {  // two channels per conversion (cvt.rn.f16x2.f32), same rounding as the scalar form
            const float a = v[g * 8 + 2 * k], b = v[g * 8 + 2 * k + 1];
            const __half2 h = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
            hw[k] = *reinterpret_cast<const uint32_t*>(&h);
            lw[k] = *reinterpret_cast<const uint32_t*>(&l);
        }
        *reinterpret_cast<uint4*>(out + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(out + plane + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
#else
        __half hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            hi[j] = __float2half_rn(v[g * 8 + j]);
            lo[j] = __float2half_rn(v[g * 8 + j] - __half2float(hi[j]));
        }
        *reinterpret_cast<uint4*>(out + off) = make_uint4(pack2h(hi[0], hi[1]), pack2h(hi[2], hi[3]), pack2h(hi[4], hi[5]), pack2h(hi[6], hi[7]));
        *reinterpret_cast<uint4*>(out + plane + off) = make_uint4(pack2h(lo[0], lo[1]), pack2h(lo[2], lo[3]), pack2h(lo[4], lo[5]), pack2h(lo[6], lo[7]));
#endif
    }
}

// Packed form of the same tensor (option "head_pack"): ONE fp16 plane of 16 channel slots per pixel -- slots 0..11 = the 12
// head channels rounded to fp16, slots 12..15 = the fp16 remainders (lo parts) of the four flow channels 8..11, whose
// weights the consuming convolution sees twice.  The flow (hundreds of pixels with sub-pixel precision) keeps its split
// hi+lo representation, the warped frames, timestep and mask (values of order 1 going into a conv) are plain fp16: 32 bytes
// per pixel instead of 64, and half the tensor-core work in the stride-2 conv that reads it.
__device__ __forceinline__ void store_c8_s2d_packed(__half* out, const float* v, int oy, int ox, int oh, int ow) {
    const size_t sub = (size_t)(oh >> 1) * (ow >> 1);
    const int par = (oy & 1) * 2 + (ox & 1);
    const size_t pix = (size_t)(oy >> 1) * (ow >> 1) + (ox >> 1);
    uint32_t g0[4], g1[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
        g0[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float a = v[8 + 2 * k], b = v[8 + 2 * k + 1];
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
        g1[k] = *reinterpret_cast<const uint32_t*>(&h);
        g1[2 + k] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(out + (((size_t)par * 2 + 0) * sub + pix) * 8) = make_uint4(g0[0], g0[1], g0[2], g0[3]);
    *reinterpret_cast<uint4*>(out + (((size_t)par * 2 + 1) * sub + pix) * 8) = make_uint4(g1[0], g1[1], g1[2], g1[3]);
}

// x0 = Interp(cat(I0, I1, T), 1/8): flownet.param:9-10
// Per-image kernel arguments are passed as __grid_constant__ structs: indexing them with blockIdx.z then reads the
// constant bank directly (a plain by-value struct is first copied to local memory, ~20 stores per thread).
struct TBatch {
    float t[V46_MAX_BATCH];
};
struct OutBatch {
    uint8_t* p[V46_MAX_BATCH];
};
struct InBatch {
    const uchar4* p0[V46_MAX_BATCH];
    const uchar4* p1[V46_MAX_BATCH];
};
struct SrcBatch {
    const uint8_t* p[2 * V46_MAX_BATCH];
};
// rife_preproc without the float conversion: RGB u8 HWC (w x h) -> RGBX [hp][wp], zeros outside the image
// bgr != 0: the frame bytes are B,G,R (the reference's Windows build, rife_preproc.comp:13,53-56); RGBX is always R,G,B,0
__global__ void rgbx_kernel(const __grid_constant__ SrcBatch sb, int w, int h, int wp, int hp, uchar4* __restrict__ out, int bgr) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= wp) return;
    uchar4 q = make_uchar4(0, 0, 0, 0);
    if (x < w && y < h) {
        const uint8_t* p = sb.p[blockIdx.z] + ((size_t)y * w + x) * 3;
        q = make_uchar4(__ldg(p + (bgr ? 2 : 0)), __ldg(p + 1), __ldg(p + (bgr ? 0 : 2)), 0);
    }
    out[((size_t)blockIdx.z * hp + y) * wp + x] = q;
}
// same, four pixels per thread: 12 source bytes as three aligned words -> one 16-byte store (w % 4 == 0, 4-byte aligned frames)
__device__ __forceinline__ uint32_t swap_rb(uint32_t v) { return (v & 0xff00ff00u) | ((v & 0xffu) << 16) | ((v >> 16) & 0xffu); }
__global__ void rgbx4_kernel(const __grid_constant__ SrcBatch sb, int w, int h, int wp, int hp, uchar4* __restrict__ out, int bgr) {
    int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x >= wp) return;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (x < w && y < h) {
        const uint32_t* p = reinterpret_cast<const uint32_t*>(sb.p[blockIdx.z] + ((size_t)y * w + x) * 3);
        const uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);  // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        o.x = a & 0x00ffffffu;
        o.y = (a >> 24) | ((b & 0xffffu) << 8);
        o.z = (b >> 16) | ((c & 0xffu) << 16);
        o.w = c >> 8;
        if (bgr) { o.x = swap_rb(o.x); o.y = swap_rb(o.y); o.z = swap_rb(o.z); o.w = swap_rb(o.w); }
    }
    *reinterpret_cast<uint4*>(out + ((size_t)blockIdx.z * hp + y) * wp + x) = o;
}

__global__ void head0_kernel(const __grid_constant__ InBatch ib, const __grid_constant__ TBatch tb, int hp, int wp, int oh, int ow, __half* __restrict__ out, int packed) {
    int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
    if (ox >= ow) return;
    const int b = blockIdx.z;
    const float t = tb.t[b];
    const Frame I0 = {ib.p0[b], wp}, I1 = {ib.p1[b], wp};
    out += (size_t)b * 16 * oh * ow * 2;
    int sx, sy;
    float fx, fy;
    lin_coeff(ox, 8.0, wp, sx, fx);
    lin_coeff(oy, 8.0, hp, sy, fy);
    const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
    float v[16];
    bilerp3(I0, sy, sx, a0, a1, b0, b1, v);
    bilerp3(I1, sy, sx, a0, a1, b0, b1, v + 3);
    v[6] = (t * a0 + t * a1) * b0 + (t * a0 + t * a1) * b1;  // the timestep plane goes through the same resize arithmetic
#pragma unroll
    for (int c = 7; c < 16; c++) v[c] = 0.f;
    if (packed) store_c8_s2d_packed(out, v, oy, ox, oh, ow);
    else store_c8_s2d_16(out, v, oy, ox, oh, ow);
}

// Block head for k >= 1 fused with the flow / mask update that follows block k-1
// (flownet.param:47-62 for k = 1, :99-115 for k = 2, :152-165 for k = 3):
//   U = bilinear(d_{k-1}, SP);  F = F + SP*U[0:4]  (k = 1: F = SP*U[0:4]);  M = M + U[4]          at full resolution
//   x = cat(bilinear(cat(warp(I0,F01), warp(I1,F23), T, M), 1/S), bilinear(F, 1/S) / S)          at 1/S resolution
// One thread owns one output pixel and its S x S full-resolution footprint: it updates (and stores) F, M for the whole
// footprint and keeps the 2x2 (S > 1) or 1 (S = 1) tap pixels the down-sampling reads -- for these integer scales the
// taps always lie inside the thread's own footprint, so no other thread's update is needed.
// MODE 0 (k = 1): F = SP*U, M = U[4]; nothing is stored and only the tap pixels are evaluated -- the next head
//                 recomputes these values from the small d_0 instead of reading 5 full-resolution planes back.
// MODE 1 (k = 2): the old F, M are recomputed as SPP*bilinear(dprev, SPP) (same operations as MODE 0), updated, stored.
// MODE 2 (k = 3): the old F, M are read from memory, updated, stored (the tail reads them).
// MODE 3 (k = 3): the old F, M are recomputed from the two coarser flow tensors (dpp at 1/8, dprev at 1/SPP), exactly
//                 as blocks 1 and 2 built them: old = (8*bilinear(dpp, 8))*1 + SPP*bilinear(dprev, SPP).
// STORE = false (MODE 1 / 3, option recompute_fm): nothing is written back -- the full-resolution F / M planes
// (20 B per pixel written here, 20 read by the next consumer) never exist; the next kernel recomputes them from the
// small d_k tensors, which stay in L2.  The arithmetic (operation order, rounding) is the same either way.
template <int S, int SP, int MODE, int SPP, bool STORE>
__global__ void head_update_kernel(const __grid_constant__ InBatch ib, float* __restrict__ F, float* __restrict__ M, const float* __restrict__ d, int dh, int dw,
                                   const float* __restrict__ dprev, int pdh, int pdw, const float* __restrict__ dpp, int ppdh, int ppdw,
                                   const __grid_constant__ TBatch tb, int hp, int wp, int oh, int ow, __half* __restrict__ out, int dch, int packed) {
    // dch = planes per image of the block outputs d / dprev / dpp: 6 for rife-v4.6 (PixelShuffle of 24 channels, the sixth
    // plane unused), 5 for rife-v4 (the deconvolution's own 5 channels at half the block resolution, so SP = 2 * S_{k-1})
    constexpr bool FIRST = MODE == 0;
    static_assert(!(STORE && MODE == 0), "MODE 0 never stores");
    int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
    if (ox >= ow) return;
    const size_t plane = (size_t)hp * wp, dplane = (size_t)dh * dw;
    const int b = blockIdx.z;
    const float t = tb.t[b];
    const Frame I0 = {ib.p0[b], wp}, I1 = {ib.p1[b], wp};
    F += (size_t)b * 4 * plane;
    M += (size_t)b * plane;
    d += (size_t)b * dch * dplane;
    const size_t pdplane = (size_t)pdh * pdw;
    if (MODE == 1 || MODE == 3) dprev += (size_t)b * dch * pdplane;
    const size_t ppdplane = (size_t)ppdh * ppdw;
    if (MODE == 3) dpp += (size_t)b * dch * ppdplane;
    out += (size_t)b * 16 * oh * ow * 2;
    constexpr int T0 = S == 1 ? 0 : S / 2 - 1;  // first tap inside the footprint (S = 4: 1, S = 2: 0)
    constexpr int NT = S == 1 ? 1 : 2;
    float tapF[NT][NT][4], tapM[NT][NT];
    // horizontal up-sampling coefficients of the S footprint columns (shared by all rows)
    int usx[S];
    float ua0[S], ua1[S];
#pragma unroll
    for (int fx = 0; fx < S; fx++) {
        float f;
        lin_coeff(S * ox + fx, 1.0 / SP, dw, usx[fx], f);
        ua0[fx] = 1.f - f;
        ua1[fx] = f;
    }
    int psx[S];
    float pa0[S], pa1[S];
    if (MODE == 1 || MODE == 3) {
#pragma unroll
        for (int fx = 0; fx < S; fx++) {
            float f;
            lin_coeff(S * ox + fx, 1.0 / SPP, pdw, psx[fx], f);
            pa0[fx] = 1.f - f;
            pa1[fx] = f;
        }
    }
    int qsx[S];
    float qa0[S], qa1[S];
    if (MODE == 3) {
#pragma unroll
        for (int fx = 0; fx < S; fx++) {
            float f;
            lin_coeff(S * ox + fx, 1.0 / 8, ppdw, qsx[fx], f);
            qa0[fx] = 1.f - f;
            qa1[fx] = f;
        }
    }
#pragma unroll
    for (int fy = 0; fy < S; fy++) {
        if (!STORE && !(fy >= T0 && fy < T0 + NT)) continue;  // nothing stored: only the tap rows matter
        const int y = S * oy + fy;
        int usy;
        float f;
        lin_coeff(y, 1.0 / SP, dh, usy, f);
        const float ub0 = 1.f - f, ub1 = f;
        float nf[4][S], nm[S];
        const size_t p0 = (size_t)y * wp + S * ox;  // first footprint pixel of this row: S consecutive floats per plane
        // vectorised row access: S floats = one 16-byte (S = 4) / 8-byte (S = 2) / 4-byte transaction per plane
        auto load_row = [&](const float* base, float* dst) {
            if constexpr (S == 4) { float4 q = *reinterpret_cast<const float4*>(base + p0); dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w; }
            else if constexpr (S == 2) { float2 q = *reinterpret_cast<const float2*>(base + p0); dst[0] = q.x; dst[1] = q.y; }
            else dst[0] = base[p0];
        };
        auto store_row = [&](float* base, const float* src) {
            if constexpr (S == 4) *reinterpret_cast<float4*>(base + p0) = make_float4(src[0], src[1], src[2], src[3]);
            else if constexpr (S == 2) *reinterpret_cast<float2*>(base + p0) = make_float2(src[0], src[1]);
            else base[p0] = src[0];
        };
        float oldf[4][S], oldm[S];
        if (MODE == 2) {
#pragma unroll
            for (int c = 0; c < 4; c++) load_row(F + c * plane, oldf[c]);
            load_row(M, oldm);
        } else if (MODE == 1) {
            int psy;
            float pf;
            lin_coeff(y, 1.0 / SPP, pdh, psy, pf);
            const float pb0 = 1.f - pf, pb1 = pf;
#pragma unroll
            for (int fx = 0; fx < S; fx++) {
#pragma unroll
                for (int c = 0; c < 4; c++) oldf[c][fx] = bilerp(dprev + c * pdplane, pdw, psy, psx[fx], pa0[fx], pa1[fx], pb0, pb1) * (float)SPP;
                oldm[fx] = bilerp(dprev + 4 * pdplane, pdw, psy, psx[fx], pa0[fx], pa1[fx], pb0, pb1);
            }
        } else if (MODE == 3) {
            int psy, qsy;
            float pf, qf;
            lin_coeff(y, 1.0 / SPP, pdh, psy, pf);
            lin_coeff(y, 1.0 / 8, ppdh, qsy, qf);
            const float pb0 = 1.f - pf, pb1 = pf, qb0 = 1.f - qf, qb1 = qf;
#pragma unroll
            for (int fx = 0; fx < S; fx++) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float f1 = bilerp(dpp + c * ppdplane, ppdw, qsy, qsx[fx], qa0[fx], qa1[fx], qb0, qb1) * 8.f;          // after block 0
                    oldf[c][fx] = f1 * 1.f + bilerp(dprev + c * pdplane, pdw, psy, psx[fx], pa0[fx], pa1[fx], pb0, pb1) * (float)SPP;  // after block 1
                }
                const float m1 = bilerp(dpp + 4 * ppdplane, ppdw, qsy, qsx[fx], qa0[fx], qa1[fx], qb0, qb1);
                oldm[fx] = m1 + bilerp(dprev + 4 * pdplane, pdw, psy, psx[fx], pa0[fx], pa1[fx], pb0, pb1);
            }
        }
#pragma unroll
        for (int fx = 0; fx < S; fx++) {
            if (!STORE && !(fx >= T0 && fx < T0 + NT)) continue;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float u = bilerp(d + c * dplane, dw, usy, usx[fx], ua0[fx], ua1[fx], ub0, ub1);
                nf[c][fx] = FIRST ? u * (float)SP : oldf[c][fx] * 1.f + u * (float)SP;  // BinaryOp mul | Eltwise SUM {1, SP}
            }
            const float um = bilerp(d + 4 * dplane, dw, usy, usx[fx], ua0[fx], ua1[fx], ub0, ub1);
            nm[fx] = FIRST ? um : oldm[fx] + um;
        }
        if (STORE) {
#pragma unroll
            for (int c = 0; c < 4; c++) store_row(F + c * plane, nf[c]);
            store_row(M, nm);
        }
#pragma unroll
        for (int fx = 0; fx < S; fx++) {
            if (fy >= T0 && fy < T0 + NT && fx >= T0 && fx < T0 + NT) {
#pragma unroll
                for (int c = 0; c < 4; c++) tapF[fy - T0][fx - T0][c] = nf[c][fx];
                tapM[fy - T0][fx - T0] = nm[fx];
            }
        }
    }
    // 8-channel vector cat(W0, W1, T, M) and the 4 flow channels at one full-resolution tap pixel
    auto at = [&](int ty, int tx, float* e) {
        const int y = S * oy + T0 + ty, x = S * ox + T0 + tx;
        const float f0 = tapF[ty][tx][0], f1 = tapF[ty][tx][1], f2 = tapF[ty][tx][2], f3 = tapF[ty][tx][3];
        WarpTap t0 = warp_tap(x, y, f0, f1, wp, hp);
        WarpTap t1 = warp_tap(x, y, f2, f3, wp, hp);
        warp_sample3(I0, t0, e);
        warp_sample3(I1, t1, e + 3);
        e[6] = t;
        e[7] = tapM[ty][tx];
        e[8] = f0; e[9] = f1; e[10] = f2; e[11] = f3;
    };
    float v[16];
    if (S == 1) {
        // Interp at identical size returns its input untouched (interp.cpp), then Concat
        float e[12];
        at(0, 0, e);
#pragma unroll
        for (int c = 0; c < 12; c++) v[c] = e[c];
    } else {
        int sx, sy;
        float fx, fy;
        lin_coeff(ox, (double)S, wp, sx, fx);  // == S*ox + T0 for these scales; kept for the exact coefficient arithmetic
        lin_coeff(oy, (double)S, hp, sy, fy);
        const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
        float e00[12], e01[12], e10[12], e11[12];
        at(0, 0, e00);
        at(0, NT - 1, e01);
        at(NT - 1, 0, e10);
        at(NT - 1, NT - 1, e11);
#pragma unroll
        for (int c = 0; c < 12; c++) {
            float row0 = e00[c] * a0 + e01[c] * a1;
            float row1 = e10[c] * a0 + e11[c] * a1;
            v[c] = row0 * b0 + row1 * b1;
        }
#pragma unroll
        for (int c = 8; c < 12; c++) v[c] = v[c] / (float)S;  // BinaryOp div by the scale (flownet.param div_17 / div_37)
    }
#pragma unroll
    for (int c = 12; c < 16; c++) v[c] = 0.f;
    if (packed) store_c8_s2d_packed(out, v, oy, ox, oh, ow);
    else store_c8_s2d_16(out, v, oy, ox, oh, ow);
}

// last update + blend + rife_postproc: flownet.param:202-217, src/rife.cpp:4375-4398, mat_pixel.cpp:158
// RC = 0: F, M (after block 2) come from memory.  RC = 1: memory holds F, M after block 1 (the head of block 3 did not
// store); the block-2 update 2*bilinear(d2, 2) is redone here.  RC = 2: nothing was ever stored, F and M are rebuilt from
// d0 (1/8), d1 (1/4) and d2 (1/2) in the order the blocks accumulated them (see head_update_kernel MODE 3).
struct DSrc {
    const float* d[3];  // d0, d1, d2: 6 planes each at 1/8, 1/4, 1/2 resolution
};
template <int SP>
__device__ __forceinline__ void up5(const float* __restrict__ d, int dh, int dw, int X, int Y, float* u) {
    int sx, sy;
    float fx, fy;
    lin_coeff(X, 1.0 / SP, dw, sx, fx);
    lin_coeff(Y, 1.0 / SP, dh, sy, fy);
    const size_t dpl = (size_t)dh * dw;
#pragma unroll
    for (int c = 0; c < 5; c++) u[c] = bilerp(d + c * dpl, dw, sy, sx, 1.f - fx, fx, 1.f - fy, fy);
}
// D3S = 1: d3 is at full resolution (rife-v4.6: flownet.param:202-207, plain adds).  D3S = 2: d3 is the 5-channel deconvolution
// output at half resolution (rife-v4: models/rife-v4/flownet.param:152-160): U = bilinear(d3, 2), F3 = F*1 + U*2, M3 = M + U[4].
template <int RC, int D3S>
__global__ void tail_kernel(const __grid_constant__ InBatch ib, const float* __restrict__ F, const float* __restrict__ M, const float* __restrict__ d3, int hp, int wp,
                            const __grid_constant__ OutBatch ob,
                            int w, int h, const __grid_constant__ DSrc ds, int contig, int dch, int bgr) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int b = blockIdx.z;
    {
        const size_t pl = (size_t)hp * wp;
        F += (size_t)b * 4 * pl; M += (size_t)b * pl; d3 += (size_t)b * dch * (pl / (D3S * D3S));
    }
    const Frame I0 = {ib.p0[b], wp}, I1 = {ib.p1[b], wp};
    uint8_t* __restrict__ rgb = ob.p[b];
    // Crop of the padded output: output pixel (y, x) is padded pixel (y, x), as rife_postproc.comp:42 does (gy * p.w + gx on
    // the padded width).  contig != 0 (option "cpu_crop_quirk") reproduces the reference's CPU path instead, which reads the
    // first w*h floats of each padded channel contiguously (rife.cpp:4375-4387): a sheared frame whenever w % 32 != 0.
    const uint32_t idx = (uint32_t)y * (uint32_t)w + (uint32_t)x;  // < 2^31: w * h pixels of one frame
    const bool same = w == wp || !contig;
    const int Y = same ? y : (int)(idx / (uint32_t)wp), X = same ? x : (int)(idx - (uint32_t)Y * (uint32_t)wp);
    const size_t plane = (size_t)hp * wp, pi = (size_t)Y * wp + X;
    float fo[4], mo;  // F, M after block 2
    if (RC == 2) {
        float u[5];
        static_assert(RC == 0 || D3S == 1, "the recompute variants exist for the rife-v4.6 layout only");
        up5<8>(ds.d[0] + (size_t)b * 6 * (hp / 8) * (wp / 8), hp / 8, wp / 8, X, Y, u);
#pragma unroll
        for (int c = 0; c < 4; c++) fo[c] = u[c] * 8.f;
        mo = u[4];
        up5<4>(ds.d[1] + (size_t)b * 6 * (hp / 4) * (wp / 4), hp / 4, wp / 4, X, Y, u);
#pragma unroll
        for (int c = 0; c < 4; c++) fo[c] = fo[c] * 1.f + u[c] * 4.f;
        mo = mo + u[4];
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++) fo[c] = F[c * plane + pi];
        mo = M[pi];
    }
    if (RC != 0) {
        float u[5];
        up5<2>(ds.d[2] + (size_t)b * 6 * (hp / 2) * (wp / 2), hp / 2, wp / 2, X, Y, u);
#pragma unroll
        for (int c = 0; c < 4; c++) fo[c] = fo[c] * 1.f + u[c] * 2.f;
        mo = mo + u[4];
    }
    float f0, f1, f2, f3, m;
    if (D3S == 1) {
        f0 = fo[0] + d3[pi]; f1 = fo[1] + d3[plane + pi];
        f2 = fo[2] + d3[2 * plane + pi]; f3 = fo[3] + d3[3 * plane + pi];
        m = mo + d3[4 * plane + pi];
    } else {
        float u[5];
        up5<D3S>(d3, hp / D3S, wp / D3S, X, Y, u);
        f0 = fo[0] * 1.f + u[0] * (float)D3S; f1 = fo[1] * 1.f + u[1] * (float)D3S;  // Eltwise SUM {1, 2}
        f2 = fo[2] * 1.f + u[2] * (float)D3S; f3 = fo[3] * 1.f + u[3] * (float)D3S;
        m = mo + u[4];
    }
    m = fminf(m, 88.3762626647949f);
    m = fmaxf(m, -88.3762626647949f);
    m = 1.f / (1.f + expf(-m));   // sigmoid.cpp:42-44
    const float om = 1.f - m;     // BinaryOp rsub
    WarpTap t0 = warp_tap(X, Y, f0, f1, wp, hp);
    WarpTap t1 = warp_tap(X, Y, f2, f3, wp, hp);
    float s0[3], s1[3];
    warp_sample3(I0, t0, s0);
    warp_sample3(I1, t1, s1);
    uint8_t* o = rgb + (size_t)idx * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float w0 = s0[c] * m;
        float w1 = s1[c] * om;
        float v = (w0 + w1) * 255.f + 0.5f;
        int iv = (int)v;
        o[bgr ? 2 - c : c] = (uint8_t)min(max(iv, 0), 255);
    }
}


}  // namespace fusedk
}  // namespace rife
