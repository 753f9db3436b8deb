// hbm_kernels.cu -- the RIFE-specific HBM-bound stages of the generic (planar fp32) path, written for coalesced 128-byte
// warp transactions on B200:
//   preproc   rife_preproc.comp / rife_preproc_tta.comp, CPU src/rife.cpp:4152-4211, 3253-3413   (a8)
//   postproc  rife_postproc.comp / rife_postproc_tta.comp / rife_out_tta_temporal_avg.comp, CPU rife.cpp:4060-4144, 4356-4398 (a9, a12)
//   flow_tta_avg   rife_flow_tta_avg.comp / rife_v2_.. / rife_v4_flow_tta_avg.comp, CPU rife.cpp:1541-1719, 3515-3668 (a11)
//   temporal merges rife_*_flow_tta_temporal_avg.comp, CPU rife.cpp:2269-2319, 4290-4311           (a12)
//   warp      src/warp.cpp:96-168 (Warp::forward)                                                 (a7)
// The eight TTA orientations (rife.cpp:3340-3364) are four row-major views of the W x H image and four of
// its transpose.  A thread block owns a 32 x 32 pixel tile: the row-major views are read / written directly with a warp
// along x, the transposed ones go through a padded shared-memory tile so that the warp runs along y in memory -- every
// global transaction is a full, aligned 128-byte line either way.  Arithmetic (operation order included) is the
// reference's; the tests compare against the oracle bit-for-bit on the u8 result (+-1 LSB allowed).
#include "kernels.h"

#include <math.h>
#include <stdio.h>

namespace rife {

static inline unsigned int cdiv(size_t a, size_t b) { return (unsigned int)((a + b - 1) / b); }

constexpr int TS = 32;  // tile side

// destination (orientation o) coordinates of padded source pixel (y, x): orientations 0-3 are [hp][wp] planes,
// 4-7 are [wp][hp] planes
__device__ __forceinline__ size_t orient_index(int o, int y, int x, int wp, int hp) {
    switch (o) {
        case 0: return (size_t)y * wp + x;
        case 1: return (size_t)y * wp + (wp - 1 - x);
        case 2: return (size_t)(hp - 1 - y) * wp + (wp - 1 - x);
        case 3: return (size_t)(hp - 1 - y) * wp + x;
        case 4: return (size_t)x * hp + y;
        case 5: return (size_t)x * hp + (hp - 1 - y);
        case 6: return (size_t)(wp - 1 - x) * hp + (hp - 1 - y);
        default: return (size_t)(wp - 1 - x) * hp + y;
    }
}

// ------------------------------------------------------------------------------------------------
// preproc: u8 HWC -> float planes * (1/255), zero outside (w, h); all `norient` orientations from ONE read of the frame.
// out: orientation o at out + o * 3 * wp * hp.  Per pixel: 3 B read, norient * 12 B written.
// ------------------------------------------------------------------------------------------------
// thread t of the 256 owns, per channel and orientation, ONE float4: four consecutive pixels along the fastest axis of that
// orientation's plane (x for the row-major views, y for the transposed ones; reversed views write the four values in reverse
// order) -- 512-byte warp stores, 3 per orientation.
__device__ __forceinline__ float4 rev4(float4 v, bool r) { return r ? make_float4(v.w, v.z, v.y, v.x) : v; }
__global__ void __launch_bounds__(256, 4) preproc_kernel(const uint8_t* __restrict__ rgb, int w, int h, float* __restrict__ out, int wp, int hp, int norient, int bgr) {
    __shared__ float tile[3][TS][TS + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int yl = ty + 8 * k, x = x0 + tx, y = y0 + yl;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (x < w && y < h) {
            const uint8_t* p = rgb + ((size_t)y * w + x) * 3;
            v0 = (float)__ldg(p + (bgr ? 2 : 0)) * (1 / 255.f);
            v1 = (float)__ldg(p + 1) * (1 / 255.f);
            v2 = (float)__ldg(p + (bgr ? 0 : 2)) * (1 / 255.f);
        }
        tile[0][yl][tx] = v0;
        tile[1][yl][tx] = v1;
        tile[2][yl][tx] = v2;
    }
    __syncthreads();
    const size_t plane = (size_t)wp * hp;
    const int a = threadIdx.x >> 3, b4 = (threadIdx.x & 7) * 4;  // slow index, first of the four fast indices
#pragma unroll 1
    for (int o = 0; o < norient; o++) {
        float* dst = out + (size_t)o * 3 * plane;
        // flips of this orientation along its fast / slow plane axis (Appendix B): o = 1, 2: x reversed; 2, 3: y reversed;
        // transposed views: 5, 6: y (fast) reversed; 6, 7: x (slow) reversed
        const bool tr = o >= 4;
        const bool rf = tr ? (o == 5 || o == 6) : (o == 1 || o == 2);
        const bool rs = tr ? (o == 6 || o == 7) : (o == 2 || o == 3);
        const int nf = tr ? hp : wp, ns = tr ? wp : hp;          // extents of the fast / slow plane axes
        const int f0 = (tr ? y0 : x0) + b4, s0 = (tr ? x0 : y0) + a;
        const size_t di = (size_t)(rs ? ns - 1 - s0 : s0) * nf + (rf ? nf - 4 - f0 : f0);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float4 v;
            if (!tr) v = make_float4(tile[c][a][b4], tile[c][a][b4 + 1], tile[c][a][b4 + 2], tile[c][a][b4 + 3]);
            else v = make_float4(tile[c][b4][a], tile[c][b4 + 1][a], tile[c][b4 + 2][a], tile[c][b4 + 3][a]);
            *reinterpret_cast<float4*>(dst + c * plane + di) = rev4(v, rf);
        }
    }
}
void launch_preproc(const uint8_t* rgb, int w, int h, float* out, int wp, int hp, int norient, int bgr, cudaStream_t st) {
    preproc_kernel<<<dim3(wp / TS, hp / TS), 256, 0, st>>>(rgb, w, h, out, wp, hp, norient, bgr);
    g_launch_count++;
}

// ------------------------------------------------------------------------------------------------
// postproc: v*255+0.5 -> (int) truncation -> clamp -> u8 (mat_pixel.cpp:158).
// n_in 1: plain.  2: temporal TTA (v + vr) * 0.5.  8 / 16: spatial TTA = mean of the 8 un-rotated outputs (/8), with
// temporal TTA the mean of the two means.
// ------------------------------------------------------------------------------------------------
struct PostArgs {
    const float* in[16];
};
__device__ __forceinline__ uint8_t quant(float v) {
    int iv = (int)v;  // truncation, as mat_pixel.cpp:158 `(uchar)min(max((int)v,0),255)`
    return (uint8_t)min(max(iv, 0), 255);
}
// no spatial TTA: four pixels per thread; 16-byte plane reads and 12-byte (3 x u32) frame writes when the row geometry allows
template <int N_IN>
__global__ void __launch_bounds__(128) postproc_plain_kernel(PostArgs pa, int wp, int hp, uint8_t* __restrict__ rgb, int w, int h, int cpu_contig, int bgr, int vec) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x >= w) return;
    const size_t plane = (size_t)wp * hp;
    // proper crop: padded pixel (y, x) (rife_postproc.comp:42).  cpu_contig: the reference CPU path's contiguous read of
    // the first w*h floats of every padded channel (rife.cpp:4375-4387); identical whenever w == wp
    const size_t idx = cpu_contig ? (size_t)y * w + x : (size_t)y * wp + x;
    float v[3][4];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        float a[4], b[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec) {
            const float4 t = *reinterpret_cast<const float4*>(pa.in[0] + q * plane + idx);
            a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
            if (N_IN == 2) { const float4 u = *reinterpret_cast<const float4*>(pa.in[1] + q * plane + idx); b[0] = u.x; b[1] = u.y; b[2] = u.z; b[3] = u.w; }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                a[j] = x + j < w ? pa.in[0][q * plane + idx + j] : 0.f;
                if (N_IN == 2) b[j] = x + j < w ? pa.in[1][q * plane + idx + j] : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) v[q][j] = N_IN == 2 ? (a[j] + b[j]) * 0.5f * 255.f + 0.5f : a[j] * 255.f + 0.5f;
    }
    uint8_t o[12];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int q = 0; q < 3; q++) o[j * 3 + (bgr ? 2 - q : q)] = quant(v[q][j]);
    uint8_t* dst = rgb + ((size_t)y * w + x) * 3;
    if (vec) {
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
        for (int k = 0; k < 3; k++) d32[k] = (uint32_t)o[4 * k] | ((uint32_t)o[4 * k + 1] << 8) | ((uint32_t)o[4 * k + 2] << 16) | ((uint32_t)o[4 * k + 3] << 24);
    } else {
        for (int j = 0; j < 4 && x + j < w; j++) { dst[j * 3] = o[j * 3]; dst[j * 3 + 1] = o[j * 3 + 1]; dst[j * 3 + 2] = o[j * 3 + 2]; }
    }
}
// spatial TTA: input i is orientation i & 7 of the padded frame (set i >> 3: forward / time-reversed).  Same float4-per-thread
// scheme as preproc_kernel: row-major inputs are read in place, transposed ones through a padded shared-memory tile.
template <int NSET>
__global__ void __launch_bounds__(256, 3) postproc_tta_kernel(PostArgs pa, int wp, int hp, uint8_t* __restrict__ rgb, int w, int h, int bgr, int vec) {
    __shared__ float T[4][TS][TS + 1];  // the transposed orientations 4-7 of one channel: T[o - 4][x][y]
    const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
    const size_t plane = (size_t)wp * hp;
    const int a = threadIdx.x >> 3, b4 = (threadIdx.x & 7) * 4;
    float acc[3][4];
#pragma unroll 1
    for (int q = 0; q < 3; q++) {
        float mean[NSET][4];
#pragma unroll
        for (int s = 0; s < NSET; s++) {
            // all eight 16-byte loads of this (channel, set) are issued before anything waits: the row-major four stay in
            // registers across the barrier (four loads in flight per thread left the kernel latency-bound at 37 % occupancy)
            float4 rv[4], tv[4];
#pragma unroll
            for (int o = 0; o < 4; o++) {  // y = y0 + a, x = x0 + b4 .. + 3
                const bool rf = o == 1 || o == 2, rs = o == 2 || o == 3;
                const size_t si = (size_t)(rs ? hp - 1 - (y0 + a) : y0 + a) * wp + (rf ? wp - 4 - (x0 + b4) : x0 + b4);
                rv[o] = *reinterpret_cast<const float4*>(pa.in[s * 8 + o] + q * plane + si);
            }
#pragma unroll
            for (int o = 4; o < 8; o++) {  // x = x0 + a (slow axis of the [wp][hp] plane), y = y0 + b4 .. + 3 (fast)
                const bool rf = o == 5 || o == 6, rs = o == 6 || o == 7;
                const size_t si = (size_t)(rs ? wp - 1 - (x0 + a) : x0 + a) * hp + (rf ? hp - 4 - (y0 + b4) : y0 + b4);
                tv[o - 4] = *reinterpret_cast<const float4*>(pa.in[s * 8 + o] + q * plane + si);
            }
            __syncthreads();
#pragma unroll
            for (int o = 4; o < 8; o++) {
                const float4 v = rev4(tv[o - 4], o == 5 || o == 6);
                T[o - 4][a][b4] = v.x; T[o - 4][a][b4 + 1] = v.y; T[o - 4][a][b4 + 2] = v.z; T[o - 4][a][b4 + 3] = v.w;
            }
            __syncthreads();
            float sum[4] = {0.f, 0.f, 0.f, 0.f};  // rife.cpp:4060-4144: the eight values are added in orientation order, then / 8
#pragma unroll
            for (int o = 0; o < 4; o++) {
                const float4 v = rev4(rv[o], o == 1 || o == 2);
                sum[0] += v.x; sum[1] += v.y; sum[2] += v.z; sum[3] += v.w;
            }
#pragma unroll
            for (int o = 4; o < 8; o++)
#pragma unroll
                for (int j = 0; j < 4; j++) sum[j] += T[o - 4][b4 + j][a];
#pragma unroll
            for (int j = 0; j < 4; j++) mean[s][j] = sum[j] / 8;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) acc[q][j] = NSET == 2 ? (mean[0][j] + mean[NSET - 1][j]) * 0.5f * 255.f + 0.5f : mean[0][j] * 255.f + 0.5f;
    }
    const int x = x0 + b4, y = y0 + a;
    if (x >= w || y >= h) return;
    uint8_t o8[12];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int q = 0; q < 3; q++) o8[j * 3 + (bgr ? 2 - q : q)] = quant(acc[q][j]);
    uint8_t* dst = rgb + ((size_t)y * w + x) * 3;
    if (vec) {  // w % 4 == 0 and a 4-byte aligned frame: the thread's 12 bytes are three aligned words
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
        for (int k = 0; k < 3; k++) d32[k] = (uint32_t)o8[4 * k] | ((uint32_t)o8[4 * k + 1] << 8) | ((uint32_t)o8[4 * k + 2] << 16) | ((uint32_t)o8[4 * k + 3] << 24);
    } else {
        for (int j = 0; j < 4 && x + j < w; j++) { dst[j * 3] = o8[j * 3]; dst[j * 3 + 1] = o8[j * 3 + 1]; dst[j * 3 + 2] = o8[j * 3 + 2]; }
    }
}
void launch_postproc(const float* const* ins, int n_in, int wp, int hp, uint8_t* rgb, int w, int h, int cpu_contig, int bgr, cudaStream_t st) {
    PostArgs pa;
    for (int i = 0; i < 16; i++) pa.in[i] = i < n_in ? ins[i] : nullptr;
    if (n_in <= 2) {
        // 16-byte reads need idx % 4 == 0 for every row; u32 writes need (y*w + x) * 3 % 4 == 0: both hold iff w % 4 == 0
        // (wp is a multiple of 32) and the frame pointer is 4-byte aligned
        const int vec = (w % 4 == 0) && (((uintptr_t)rgb & 3) == 0);
        dim3 g(cdiv((size_t)(w + 3) / 4, 128), h);
        if (n_in == 1) postproc_plain_kernel<1><<<g, 128, 0, st>>>(pa, wp, hp, rgb, w, h, cpu_contig && w != wp, bgr, vec);
        else postproc_plain_kernel<2><<<g, 128, 0, st>>>(pa, wp, hp, rgb, w, h, cpu_contig && w != wp, bgr, vec);
    } else {
        const int vec = (w % 4 == 0) && (((uintptr_t)rgb & 3) == 0);
        dim3 g(wp / TS, hp / TS);
        if (n_in == 8) postproc_tta_kernel<1><<<g, 256, 0, st>>>(pa, wp, hp, rgb, w, h, bgr, vec);
        else postproc_tta_kernel<2><<<g, 256, 0, st>>>(pa, wp, hp, rgb, w, h, bgr, vec);
    }
    g_launch_count++;
}

// ------------------------------------------------------------------------------------------------
// temporal merges (in place), n = elements per channel; float4 per thread when n % 4 == 0
// ------------------------------------------------------------------------------------------------
// rife.cpp:2307-2319 (v1 rule) -- rife_flow_tta_temporal_avg.comp:19-42
template <typename V>
__global__ void temporal_merge_v1_kernel(float* __restrict__ f, float* __restrict__ fr, size_t n) {
    constexpr int L = sizeof(V) / 4;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * L;
    if (i >= n) return;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        V a = *reinterpret_cast<const V*>(f + c * n + i), b = *reinterpret_cast<const V*>(fr + c * n + i);
        float* pa = reinterpret_cast<float*>(&a);
        float* pb = reinterpret_cast<float*>(&b);
#pragma unroll
        for (int j = 0; j < L; j++) { const float x = (pa[j] - pb[j]) * 0.5f; pa[j] = x; pb[j] = -x; }
        *reinterpret_cast<V*>(f + c * n + i) = a;
        *reinterpret_cast<V*>(fr + c * n + i) = b;
    }
}
void launch_temporal_merge_v1(float* f, float* fr, size_t n, cudaStream_t st) {
    if (n % 4 == 0) temporal_merge_v1_kernel<float4><<<cdiv(n / 4, 256), 256, 0, st>>>(f, fr, n);
    else temporal_merge_v1_kernel<float><<<cdiv(n, 256), 256, 0, st>>>(f, fr, n);
    g_launch_count++;
}
// rife.cpp:2285-2306 (v2 rule), :4290-4311 (v4 adds the mask) -- rife_v2/v4_flow_tta_temporal_avg.comp
template <typename V>
__global__ void temporal_merge_v2_kernel(float* __restrict__ f, float* __restrict__ fr, size_t n, int has_mask) {
    constexpr int L = sizeof(V) / 4;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * L;
    if (i >= n) return;
    V F[4], R[4], a, b;
#pragma unroll
    for (int c = 0; c < 4; c++) { F[c] = *reinterpret_cast<const V*>(f + c * n + i); R[c] = *reinterpret_cast<const V*>(fr + c * n + i); }
    if (has_mask) { a = *reinterpret_cast<const V*>(f + 4 * n + i); b = *reinterpret_cast<const V*>(fr + 4 * n + i); }  // all ten loads in flight before the first store
    V X, Y, Z, W;
#pragma unroll
    for (int j = 0; j < L; j++) {
        reinterpret_cast<float*>(&X)[j] = (reinterpret_cast<float*>(&F[0])[j] + reinterpret_cast<float*>(&R[2])[j]) * 0.5f;
        reinterpret_cast<float*>(&Y)[j] = (reinterpret_cast<float*>(&F[1])[j] + reinterpret_cast<float*>(&R[3])[j]) * 0.5f;
        reinterpret_cast<float*>(&Z)[j] = (reinterpret_cast<float*>(&F[2])[j] + reinterpret_cast<float*>(&R[0])[j]) * 0.5f;
        reinterpret_cast<float*>(&W)[j] = (reinterpret_cast<float*>(&F[3])[j] + reinterpret_cast<float*>(&R[1])[j]) * 0.5f;
    }
    *reinterpret_cast<V*>(f + i) = X; *reinterpret_cast<V*>(f + n + i) = Y; *reinterpret_cast<V*>(f + 2 * n + i) = Z; *reinterpret_cast<V*>(f + 3 * n + i) = W;
    *reinterpret_cast<V*>(fr + i) = Z; *reinterpret_cast<V*>(fr + n + i) = W; *reinterpret_cast<V*>(fr + 2 * n + i) = X; *reinterpret_cast<V*>(fr + 3 * n + i) = Y;
    if (has_mask) {
#pragma unroll
        for (int j = 0; j < L; j++) {
            const float m = (reinterpret_cast<float*>(&a)[j] - reinterpret_cast<float*>(&b)[j]) * 0.5f;
            reinterpret_cast<float*>(&a)[j] = m;
            reinterpret_cast<float*>(&b)[j] = -m;
        }
        *reinterpret_cast<V*>(f + 4 * n + i) = a;
        *reinterpret_cast<V*>(fr + 4 * n + i) = b;
    }
}
void launch_temporal_merge_v2(float* f, float* fr, size_t n, int has_mask, cudaStream_t st) {
    if (n % 4 == 0) temporal_merge_v2_kernel<float4><<<cdiv(n / 4, 256), 256, 0, st>>>(f, fr, n, has_mask);
    else temporal_merge_v2_kernel<float><<<cdiv(n, 256), 256, 0, st>>>(f, fr, n, has_mask);
    g_launch_count++;
}

// ------------------------------------------------------------------------------------------------
// spatial TTA flow average over the 8 orientation blobs, in place (rife.cpp:1541-1719 v1/v2, :3515-3668 v4).
// Pair k of channels (2k, 2k+1) is an (x, y) flow; channel 4 (nch == 5) is the mask (plain mean).
// Blobs 0-3 are [nch][fh][fw], blobs 4-7 are [nch][fw][fh].  Every blob element belongs to exactly one tile, so the
// in-place update needs no ordering between blocks.
// ------------------------------------------------------------------------------------------------
struct Flow8 {
    float* f[8];
};
__global__ void __launch_bounds__(256, 3) flow_tta_avg_kernel(Flow8 F, int nch, int fw, int fh) {
    __shared__ float T[4][2][TS][TS + 1];  // transposed orientations 4-7, two channels: T[o - 4][c][x][y]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
    const size_t plane = (size_t)fw * fh;
    const int npair = nch >= 4 ? 2 : 1;
    const int ngroups = npair + (nch == 5 ? 1 : 0);
#pragma unroll 1
    for (int g = 0; g < ngroups; g++) {
        const bool mask = g == npair;
        const size_t cx = (size_t)(mask ? 4 : 2 * g) * plane, cy = (size_t)(mask ? 4 : 2 * g + 1) * plane;
        __syncthreads();
#pragma unroll
        for (int o = 4; o < 8; o++)
#pragma unroll
            for (int k = 0; k < 4; k++) {  // lane = y: contiguous in the [fw][fh] planes
                const int xl = ty + 8 * k, x = x0 + xl, y = y0 + tx;
                if (x < fw && y < fh) {
                    const size_t id = orient_index(o, y, x, fw, fh);
                    T[o - 4][0][xl][tx] = F.f[o][cx + id];
                    if (!mask) T[o - 4][1][xl][tx] = F.f[o][cy + id];
                }
            }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int yl = ty + 8 * k, x = x0 + tx, y = y0 + yl;
            if (x >= fw || y >= fh) continue;
            size_t id[4];
#pragma unroll
            for (int o = 0; o < 4; o++) id[o] = orient_index(o, y, x, fw, fh);
            if (mask) {
                float m = 0.f;
#pragma unroll
                for (int o = 0; o < 4; o++) m += F.f[o][cx + id[o]];
#pragma unroll
                for (int o = 4; o < 8; o++) m += T[o - 4][0][tx][yl];
                m *= 0.125f;
#pragma unroll
                for (int o = 0; o < 4; o++) F.f[o][cx + id[o]] = m;
#pragma unroll
                for (int o = 4; o < 8; o++) T[o - 4][0][tx][yl] = m;
            } else {
                // un-rotate and average (signs and the x/y swap of the transposed orientations: SURVEY.md section 2.3)
                const float vx = (F.f[0][cx + id[0]] + -F.f[1][cx + id[1]] + -F.f[2][cx + id[2]] + F.f[3][cx + id[3]] +
                                  T[0][1][tx][yl] + T[1][1][tx][yl] + -T[2][1][tx][yl] + -T[3][1][tx][yl]) * 0.125f;
                const float vy = (F.f[0][cy + id[0]] + F.f[1][cy + id[1]] + -F.f[2][cy + id[2]] + -F.f[3][cy + id[3]] +
                                  T[0][0][tx][yl] + -T[1][0][tx][yl] + -T[2][0][tx][yl] + T[3][0][tx][yl]) * 0.125f;
                F.f[0][cx + id[0]] = vx;  F.f[1][cx + id[1]] = -vx; F.f[2][cx + id[2]] = -vx; F.f[3][cx + id[3]] = vx;
                F.f[0][cy + id[0]] = vy;  F.f[1][cy + id[1]] = vy;  F.f[2][cy + id[2]] = -vy; F.f[3][cy + id[3]] = -vy;
                T[0][0][tx][yl] = vy;  T[1][0][tx][yl] = -vy; T[2][0][tx][yl] = -vy; T[3][0][tx][yl] = vy;
                T[0][1][tx][yl] = vx;  T[1][1][tx][yl] = vx;  T[2][1][tx][yl] = -vx; T[3][1][tx][yl] = -vx;
            }
        }
        __syncthreads();
#pragma unroll
        for (int o = 4; o < 8; o++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int xl = ty + 8 * k, x = x0 + xl, y = y0 + tx;
                if (x < fw && y < fh) {
                    const size_t id = orient_index(o, y, x, fw, fh);
                    F.f[o][cx + id] = T[o - 4][0][xl][tx];
                    if (!mask) F.f[o][cy + id] = T[o - 4][1][xl][tx];
                }
            }
    }
}
// Same computation, four consecutive elements per thread (fw % 4 == 0 and fh % 4 == 0: every float4 below is aligned and lies
// entirely inside or outside the blob): row-major blobs are read and written in place as float4 along x, the transposed ones
// as float4 along y through the shared-memory tile.
__device__ __forceinline__ void ld4(const float* p, bool rev, float* d) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    if (rev) { d[0] = v.w; d[1] = v.z; d[2] = v.y; d[3] = v.x; }
    else { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
}
__device__ __forceinline__ void st4(float* p, bool rev, const float* d, float sign) {
    *reinterpret_cast<float4*>(p) = rev ? make_float4(sign * d[3], sign * d[2], sign * d[1], sign * d[0]) : make_float4(sign * d[0], sign * d[1], sign * d[2], sign * d[3]);
}
__global__ void __launch_bounds__(256, 3) flow_tta_avg4_kernel(Flow8 F, int nch, int fw, int fh) {
    __shared__ float T[4][2][TS][TS + 1];  // transposed orientations 4-7, two channels: T[o - 4][c][x][y]
    const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
    const int a = threadIdx.x >> 3, b4 = (threadIdx.x & 7) * 4;
    const size_t plane = (size_t)fw * fh;
    const int npair = nch >= 4 ? 2 : 1;
    const int ngroups = npair + (nch == 5 ? 1 : 0);
    // element offsets of this thread's float4 in the transposed blobs (x = x0 + a slow, y = y0 + b4 .. fast) and in the
    // row-major ones (y = y0 + a slow, x = x0 + b4 .. fast); reversed axes run backwards in memory
    const bool tin = x0 + a < fw && y0 + b4 < fh, rin = y0 + a < fh && x0 + b4 < fw;
    size_t ti[4], ri[4];
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const bool trf = o == 1 || o == 2, trs = o == 2 || o == 3;  // orientation 4 + o: y (fast) reversed for 5, 6; x (slow) reversed for 6, 7
        ti[o] = (size_t)(trs ? fw - 1 - (x0 + a) : x0 + a) * fh + (trf ? fh - 4 - (y0 + b4) : y0 + b4);
        const bool rf = o == 1 || o == 2, rs = o == 2 || o == 3;    // orientation o: x reversed for 1, 2; y reversed for 2, 3
        ri[o] = (size_t)(rs ? fh - 1 - (y0 + a) : y0 + a) * fw + (rf ? fw - 4 - (x0 + b4) : x0 + b4);
    }
#pragma unroll 1
    for (int g = 0; g < ngroups; g++) {
        const bool mask = g == npair;
        const size_t cx = (size_t)(mask ? 4 : 2 * g) * plane, cy = (size_t)(mask ? 4 : 2 * g + 1) * plane;
        __syncthreads();
        if (tin) {
#pragma unroll
            for (int o = 0; o < 4; o++) {
                float v[4];
                ld4(F.f[4 + o] + cx + ti[o], o == 1 || o == 2, v);
#pragma unroll
                for (int j = 0; j < 4; j++) T[o][0][a][b4 + j] = v[j];
                if (!mask) {
                    ld4(F.f[4 + o] + cy + ti[o], o == 1 || o == 2, v);
#pragma unroll
                    for (int j = 0; j < 4; j++) T[o][1][a][b4 + j] = v[j];
                }
            }
        }
        __syncthreads();
        if (rin) {
            if (mask) {
                float m[4] = {0.f, 0.f, 0.f, 0.f}, v[4];
#pragma unroll
                for (int o = 0; o < 4; o++) {
                    ld4(F.f[o] + cx + ri[o], o == 1 || o == 2, v);
#pragma unroll
                    for (int j = 0; j < 4; j++) m[j] += v[j];
                }
#pragma unroll
                for (int o = 0; o < 4; o++)
#pragma unroll
                    for (int j = 0; j < 4; j++) m[j] += T[o][0][b4 + j][a];
#pragma unroll
                for (int j = 0; j < 4; j++) m[j] *= 0.125f;
#pragma unroll
                for (int o = 0; o < 4; o++) st4(F.f[o] + cx + ri[o], o == 1 || o == 2, m, 1.f);
#pragma unroll
                for (int o = 0; o < 4; o++)
#pragma unroll
                    for (int j = 0; j < 4; j++) T[o][0][b4 + j][a] = m[j];
            } else {
                float fx[4][4], fy[4][4], vx[4], vy[4];
#pragma unroll
                for (int o = 0; o < 4; o++) { ld4(F.f[o] + cx + ri[o], o == 1 || o == 2, fx[o]); ld4(F.f[o] + cy + ri[o], o == 1 || o == 2, fy[o]); }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // un-rotate and average (signs and the x/y swap of the transposed orientations: SURVEY.md section 2.3)
                    vx[j] = (fx[0][j] + -fx[1][j] + -fx[2][j] + fx[3][j] + T[0][1][b4 + j][a] + T[1][1][b4 + j][a] + -T[2][1][b4 + j][a] + -T[3][1][b4 + j][a]) * 0.125f;
                    vy[j] = (fy[0][j] + fy[1][j] + -fy[2][j] + -fy[3][j] + T[0][0][b4 + j][a] + -T[1][0][b4 + j][a] + -T[2][0][b4 + j][a] + T[3][0][b4 + j][a]) * 0.125f;
                }
                const float sx[4] = {1.f, -1.f, -1.f, 1.f}, sy[4] = {1.f, 1.f, -1.f, -1.f};
#pragma unroll
                for (int o = 0; o < 4; o++) { st4(F.f[o] + cx + ri[o], o == 1 || o == 2, vx, sx[o]); st4(F.f[o] + cy + ri[o], o == 1 || o == 2, vy, sy[o]); }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    T[0][0][b4 + j][a] = vy[j];  T[1][0][b4 + j][a] = -vy[j]; T[2][0][b4 + j][a] = -vy[j]; T[3][0][b4 + j][a] = vy[j];
                    T[0][1][b4 + j][a] = vx[j];  T[1][1][b4 + j][a] = vx[j];  T[2][1][b4 + j][a] = -vx[j]; T[3][1][b4 + j][a] = -vx[j];
                }
            }
        }
        __syncthreads();
        if (tin) {
#pragma unroll
            for (int o = 0; o < 4; o++) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = T[o][0][a][b4 + j];
                st4(F.f[4 + o] + cx + ti[o], o == 1 || o == 2, v, 1.f);
                if (!mask) {
#pragma unroll
                    for (int j = 0; j < 4; j++) v[j] = T[o][1][a][b4 + j];
                    st4(F.f[4 + o] + cy + ti[o], o == 1 || o == 2, v, 1.f);
                }
            }
        }
    }
}
void launch_flow_tta_avg(float* const* f8, int nch, int fw, int fh, cudaStream_t st) {
    Flow8 F;
    for (int i = 0; i < 8; i++) F.f[i] = f8[i];
    bool vec = fw % 4 == 0 && fh % 4 == 0;
    for (int i = 0; i < 8; i++) vec = vec && ((uintptr_t)f8[i] & 15) == 0;
    if (vec) flow_tta_avg4_kernel<<<dim3(cdiv(fw, TS), cdiv(fh, TS)), 256, 0, st>>>(F, nch, fw, fh);
    else flow_tta_avg_kernel<<<dim3(cdiv(fw, TS), cdiv(fh, TS)), 256, 0, st>>>(F, nch, fw, fh);
    g_launch_count++;
}

// ------------------------------------------------------------------------------------------------
// src/warp.cpp:96-168: backward bilinear warp, indices clamped, alpha / beta taken AFTER clamping.
// One thread per pixel and group of CG channels: the two flow values and the tap geometry are computed once, the 4*CG
// gathers of the group are all issued before the first use (the warp's lanes walk neighbouring pixels, so for a smooth
// flow each gather instruction touches one or two 128-byte lines); stores are full coalesced lines.
// ------------------------------------------------------------------------------------------------
// Address arithmetic was most of this kernel (43 LEA + 36 IADD3 of 240 SASS instructions for 3 channels: every gather formed
// its 64-bit address from scratch); now the four tap pointers are formed once and stepped plane by plane with one IMAD.WIDE
// each (the empty asm keeps the compiler from folding the steps back into per-load index arithmetic): 3 ch 187 -> ~115
// executed instructions per pixel.  FULL = all CG channels of the group exist (no per-channel predicate).
template <int CG, bool FULL>
__global__ void __launch_bounds__(128) warp_kernel(const float* __restrict__ img, const float* __restrict__ flow, float* __restrict__ out, int c, int h, int w) {
    const unsigned x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= (unsigned)w) return;
    const unsigned hw = (unsigned)h * (unsigned)w, pi = y * (unsigned)w + x;  // one plane holds fewer than 2^31 elements (frames up to 32k x 32k / 4)
    const float sx = (float)(int)x + __ldg(flow + pi), sy = (float)(int)y + __ldg(flow + hw + pi);
    int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), w - 1);
    y0 = min(max(y0, 0), h - 1);
    x1 = min(max(x1, 0), w - 1);
    y1 = min(max(y1, 0), h - 1);
    const float alpha = sx - x0, beta = sy - y0;
    const unsigned r0 = (unsigned)y0 * (unsigned)w, r1 = (unsigned)y1 * (unsigned)w;
    const int q0 = blockIdx.z * CG;
    const float* ib = img + (size_t)q0 * hw;
    const float* p00 = ib + (r0 + x0);
    const float* p01 = ib + (r0 + x1);
    const float* p10 = ib + (r1 + x0);
    const float* p11 = ib + (r1 + x1);
    float* po = out + (size_t)q0 * hw + pi;
    float v0[CG], v1[CG], v2[CG], v3[CG];
#pragma unroll
    for (int j = 0; j < CG; j++) {
        if (FULL || q0 + j < c) {
            v0[j] = __ldg(p00); v1[j] = __ldg(p01); v2[j] = __ldg(p10); v3[j] = __ldg(p11);
            asm volatile("" : "+l"(p00), "+l"(p01), "+l"(p10), "+l"(p11));
            p00 += hw; p01 += hw; p10 += hw; p11 += hw;
        }
    }
#pragma unroll
    for (int j = 0; j < CG; j++) {
        if (FULL || q0 + j < c) {
            const float v4 = v0[j] * (1 - alpha) + v1[j] * alpha;
            const float v5 = v2[j] * (1 - alpha) + v3[j] * alpha;
            *po = v4 * (1 - beta) + v5 * beta;
            po += hw;
        }
    }
}
void launch_warp(const float* img, const float* flow, float* out, int c, int h, int w, cudaStream_t st) {
    const dim3 g1(cdiv(w, 128), h, 1);
    if (c == 3) warp_kernel<3, true><<<g1, 128, 0, st>>>(img, flow, out, c, h, w);
    else if (c <= 4) warp_kernel<4, false><<<g1, 128, 0, st>>>(img, flow, out, c, h, w);
    else if (c % 8 == 0) warp_kernel<8, true><<<dim3(cdiv(w, 128), h, c / 8), 128, 0, st>>>(img, flow, out, c, h, w);
    else warp_kernel<8, false><<<dim3(cdiv(w, 128), h, cdiv(c, 8)), 128, 0, st>>>(img, flow, out, c, h, w);
    g_launch_count++;
}

}  // namespace rife
