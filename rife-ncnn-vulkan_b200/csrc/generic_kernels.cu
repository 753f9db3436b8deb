// generic_kernels.cu -- fp32 planar (CHW) kernels for the generic graph executor ("exact" precision tier)
// and the RIFE-specific HBM stages.  Semantics follow the reference's generic layer definitions
// (src/ncnn/src/layer/*.cpp, cited per kernel) and src/warp.cpp / src/rife_*.comp; the code is ours.
#include "kernels.h"

#include <math.h>
#include <stdio.h>

namespace rife {

unsigned long long g_launch_count = 0;
unsigned long long g_h2d_bytes = 0, g_d2h_bytes = 0;

static inline unsigned int cdiv(size_t a, size_t b) { return (unsigned int)((a + b - 1) / b); }

__device__ __forceinline__ float apply_act(float v, int act, float p0, float p1) {
    // src/ncnn/src/layer/fused_activation.h:22-75
    switch (act) {
        case 1: v = fmaxf(v, 0.f); break;
        case 2: v = v > 0.f ? v : v * p0; break;
        case 3: v = fminf(fmaxf(v, p0), p1); break;
        case 4:
            v = fminf(v, 88.3762626647949f);
            v = fmaxf(v, -88.3762626647949f);
            v = 1.f / (1.f + expf(-v));
            break;
        default: break;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// Direct convolution, fp32, CUDA cores.  Convolution semantics: src/ncnn/src/layer/convolution.cpp:133-204
// (weights [oc][ic][kh*kw], zero padding); deconvolution 4x4 s2 p1 (deconvolution.cpp:68-141, scatter form,
// no kernel flip) is evaluated in gather form as 4 output-parity classes of 2x2 taps each.
// Block = 256 threads computing a 32x16 tile of the compute domain for OCT output channels.
// ------------------------------------------------------------------------------------------------
constexpr int TW = 32, TH = 16, PXT = 2;

template <int K, int S, int OCT, int ICC>
__global__ void __launch_bounds__(256) conv_direct_kernel(ConvArgs a) {
    constexpr int KK = K * K;
    constexpr int IW_T = (TW - 1) * S + K;
    constexpr int IH_T = (TH - 1) * S + K;
    constexpr int IW_P = IW_T | 1;  // odd row pitch: fewer bank conflicts for S == 2
    extern __shared__ float smem[];
    float* in_s = smem;                       // [ICC][IH_T][IW_P]
    float* w_s = smem + ICC * IH_T * IW_P;    // [ICC][KK][OCT]

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int tiles_x = (a.DW + TW - 1) / TW;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int octiles = (a.Cout + OCT - 1) / OCT;
    const int parity = blockIdx.y / octiles;
    const int oc0 = (blockIdx.y % octiles) * OCT;
    const int py = a.nparity > 1 ? (parity >> 1) : 0, px = a.nparity > 1 ? (parity & 1) : 0;

    const int dx0 = tile_x * TW, dy0 = tile_y * TH;
    const int ix0 = dx0 * S + a.in_off_x + px, iy0 = dy0 * S + a.in_off_y + py;
    const float* wT = a.wT + (size_t)parity * a.Cin * KK * a.ocpad;

    float acc[PXT][OCT];
#pragma unroll
    for (int p = 0; p < PXT; p++)
#pragma unroll
        for (int o = 0; o < OCT; o++) acc[p][o] = 0.f;

    for (int ic0 = 0; ic0 < a.Cin; ic0 += ICC) {
        __syncthreads();
        // stage the input patch (zero outside the image / beyond Cin)
        for (int idx = threadIdx.x; idx < ICC * IH_T * IW_T; idx += 256) {
            int c = idx / (IH_T * IW_T), r = idx % (IH_T * IW_T);
            int yy = r / IW_T, xx = r % IW_T;
            int gy = iy0 + yy, gx = ix0 + xx, gc = ic0 + c;
            float v = 0.f;
            if (gc < a.Cin && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) v = __ldg(a.in + ((size_t)gc * a.H + gy) * a.W + gx);
            in_s[(c * IH_T + yy) * IW_P + xx] = v;
        }
        for (int idx = threadIdx.x; idx < ICC * KK * OCT; idx += 256) {
            int c = idx / (KK * OCT), r = idx % (KK * OCT);
            int kk = r / OCT, o = r % OCT;
            int gc = ic0 + c;
            w_s[idx] = gc < a.Cin ? __ldg(wT + ((size_t)gc * KK + kk) * a.ocpad + oc0 + o) : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < ICC; c++) {
#pragma unroll
            for (int ky = 0; ky < K; ky++) {
#pragma unroll
                for (int kx = 0; kx < K; kx++) {
                    float v0 = in_s[(c * IH_T + ty * S + ky) * IW_P + tx * S + kx];
                    float v1 = in_s[(c * IH_T + (ty + 8) * S + ky) * IW_P + tx * S + kx];
                    const float4* wp = reinterpret_cast<const float4*>(w_s + (c * KK + ky * K + kx) * OCT);
#pragma unroll
                    for (int o4 = 0; o4 < OCT / 4; o4++) {
                        float4 w = wp[o4];
                        acc[0][o4 * 4 + 0] = fmaf(v0, w.x, acc[0][o4 * 4 + 0]);
                        acc[0][o4 * 4 + 1] = fmaf(v0, w.y, acc[0][o4 * 4 + 1]);
                        acc[0][o4 * 4 + 2] = fmaf(v0, w.z, acc[0][o4 * 4 + 2]);
                        acc[0][o4 * 4 + 3] = fmaf(v0, w.w, acc[0][o4 * 4 + 3]);
                        acc[1][o4 * 4 + 0] = fmaf(v1, w.x, acc[1][o4 * 4 + 0]);
                        acc[1][o4 * 4 + 1] = fmaf(v1, w.y, acc[1][o4 * 4 + 1]);
                        acc[1][o4 * 4 + 2] = fmaf(v1, w.z, acc[1][o4 * 4 + 2]);
                        acc[1][o4 * 4 + 3] = fmaf(v1, w.w, acc[1][o4 * 4 + 3]);
                    }
                }
            }
        }
    }

#pragma unroll
    for (int p = 0; p < PXT; p++) {
        int dy = dy0 + ty + p * 8, dx = dx0 + tx;
        if (dy >= a.DH || dx >= a.DW) continue;
        int oy = dy * a.out_mul + a.out_off_y + py, ox = dx * a.out_mul + a.out_off_x + px;
        if (oy >= a.OH || ox >= a.OW) continue;
#pragma unroll
        for (int o = 0; o < OCT; o++) {
            int oc = oc0 + o;
            if (oc >= a.Cout) break;
            float v = acc[p][o] + (a.bias ? __ldg(a.bias + oc) : 0.f);
            v = apply_act(v, a.act, a.act_p0, a.act_p1);
            size_t oi = ((size_t)oc * a.OH + oy) * a.OW + ox;
            if (a.res) v += __ldg(a.res + oi);
            if (a.post_act == 2) v = v > 0.f ? v : v * a.post_p0;
            else if (a.post_act == 5) { float s = __ldg(a.post_slope + oc); v = v < 0.f ? v * s : v; }
            a.out[oi] = v;
        }
    }
}

template <int K, int S, int OCT, int ICC>
static void launch_conv_t(const ConvArgs& a, cudaStream_t st) {
    constexpr int IW_T = (TW - 1) * S + K, IH_T = (TH - 1) * S + K, IW_P = IW_T | 1;
    size_t smem = sizeof(float) * (ICC * IH_T * IW_P + ICC * K * K * OCT);
    // the attribute is per device: one process may drive several GPUs (src/main.cpp -g 0,1,...)
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(conv_direct_kernel<K, S, OCT, ICC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[dev] = true;
    }
    int tiles = ((a.DW + TW - 1) / TW) * ((a.DH + TH - 1) / TH);
    int octiles = (a.Cout + OCT - 1) / OCT;  // weights are padded to ocpad (multiple of 64), so partial tiles read zeros
    conv_direct_kernel<K, S, OCT, ICC><<<dim3(tiles, octiles * a.nparity), 256, smem, st>>>(a);
    g_launch_count++;
}

template <int K, int S>
static void launch_conv_ks(const ConvArgs& a, cudaStream_t st) {
    // pick the oc tile: wide for real layers, narrow for the small flow / mask heads
    if (a.Cout > 32) launch_conv_t<K, S, 64, (K == 5 ? 4 : 8)>(a, st);
    else if (a.Cout > 16) launch_conv_t<K, S, 32, (K == 5 ? 4 : 8)>(a, st);
    else launch_conv_t<K, S, 16, (K == 5 ? 4 : 8)>(a, st);
}

void launch_conv(const ConvArgs& a, int K, int S, cudaStream_t st) {
    if (K == 3 && S == 1) launch_conv_ks<3, 1>(a, st);
    else if (K == 3 && S == 2) launch_conv_ks<3, 2>(a, st);
    else if (K == 5 && S == 1) launch_conv_ks<5, 1>(a, st);
    else if (K == 5 && S == 2) launch_conv_ks<5, 2>(a, st);
    else if (K == 2 && S == 1) launch_conv_ks<2, 1>(a, st);
    else if (K == 1 && S == 1) launch_conv_ks<1, 1>(a, st);
    else fprintf(stderr, "rife_b200: unsupported conv K=%d S=%d\n", K, S);
}

// ------------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------------
__global__ void unary_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, int op, float p0, float p1) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float v = in[i];
        switch (op) {
            case U_RELU: v = fmaxf(v, 0.f); break;                       // relu.cpp:27-66
            case U_LEAKY: v = v < 0.f ? v * p0 : v; break;
            case U_SIGMOID:                                              // sigmoid.cpp:42-44
                v = fminf(v, 88.3762626647949f);
                v = fmaxf(v, -88.3762626647949f);
                v = 1.f / (1.f + expf(-v));
                break;
            case U_CLIP: v = fminf(fmaxf(v, p0), p1); break;            // clip.cpp
            case U_NEG: v = -v; break;                                   // unaryop.h:34
            case U_ADD_S: v = v + p0; break;                             // binaryop.cpp scalar forms
            case U_SUB_S: v = v - p0; break;
            case U_MUL_S: v = v * p0; break;
            case U_DIV_S: v = v / p0; break;
            case U_RSUB_S: v = p0 - v; break;
            case U_RDIV_S: v = p0 / v; break;
            default: break;
        }
        out[i] = v;
    }
}
void launch_unary(const float* in, float* out, size_t n, int op, float p0, float p1, cudaStream_t st) {
    if (!n) return;
    unsigned int blocks = min(cdiv(n, 256), 148u * 16);
    unary_kernel<<<blocks, 256, 0, st>>>(in, out, n, op, p0, p1);
    g_launch_count++;
}

// prelu.cpp:27-110 (per-channel slope, or a single shared slope)
__global__ void prelu_kernel(const float* __restrict__ in, const float* __restrict__ slope, int nslope, float* __restrict__ out, int c, size_t hw) {
    size_t n = (size_t)c * hw;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float v = in[i];
        float s = nslope > 1 ? __ldg(slope + i / hw) : __ldg(slope);
        out[i] = v < 0.f ? v * s : v;
    }
}
void launch_prelu(const float* in, const float* slope, int nslope, float* out, int c, size_t hw, cudaStream_t st) {
    size_t n = (size_t)c * hw;
    prelu_kernel<<<min(cdiv(n, 256), 148u * 16), 256, 0, st>>>(in, slope, nslope, out, c, hw);
    g_launch_count++;
}

// binaryop.cpp:60-330 broadcasting subset used by the models: full x full, x per-channel, x single-plane
__global__ void binary_kernel(const float* __restrict__ a, int ac, size_t ahw, const float* __restrict__ b, int bc, size_t bhw,
                              float* __restrict__ out, int c, size_t hw, int op) {
    size_t n = (size_t)c * hw;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        size_t q = i / hw, r = i - q * hw;
        float x = a[(ac == 1 ? 0 : q) * ahw + (ahw == 1 ? 0 : r)];
        float y = b[(bc == 1 ? 0 : q) * bhw + (bhw == 1 ? 0 : r)];
        float v;
        switch (op) {
            case B_ADD: v = x + y; break;
            case B_SUB: v = x - y; break;
            case B_MUL: v = x * y; break;
            case B_DIV: v = x / y; break;
            case B_MAX: v = fmaxf(x, y); break;
            case B_MIN: v = fminf(x, y); break;
            case B_RSUB: v = y - x; break;
            case B_RDIV: v = y / x; break;
            default: v = powf(x, y); break;
        }
        out[i] = v;
    }
}
// the two shapes the models use almost exclusively: full x full and full x per-channel scalar (the SE scale): one grid row per
// channel, no per-element division, float4 when the plane allows
template <bool BCAST>
__global__ void binary_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t hw, int op, int vec) {
    const size_t q = blockIdx.y;
    const float* pa = a + q * hw;
    const float* pb = BCAST ? b + q : b + q * hw;
    float* po = out + q * hw;
    const float ys = BCAST ? __ldg(pb) : 0.f;
    auto f = [op](float x, float y) -> float {
        switch (op) {
            case B_ADD: return x + y;
            case B_SUB: return x - y;
            case B_MUL: return x * y;
            case B_DIV: return x / y;
            case B_MAX: return fmaxf(x, y);
            case B_MIN: return fminf(x, y);
            case B_RSUB: return y - x;
            case B_RDIV: return y / x;
            default: return powf(x, y);
        }
    };
    if (vec) {
        const size_t n4 = hw / 4;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 x = reinterpret_cast<const float4*>(pa)[i];
            float4 y = make_float4(ys, ys, ys, ys);
            if (!BCAST) y = reinterpret_cast<const float4*>(pb)[i];
            reinterpret_cast<float4*>(po)[i] = make_float4(f(x.x, y.x), f(x.y, y.y), f(x.z, y.z), f(x.w, y.w));
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) po[i] = f(pa[i], BCAST ? ys : pb[i]);
    }
}
void launch_binary(const float* a, int ac, size_t ahw, const float* b, int bc, size_t bhw, float* out, int c, size_t hw, int op, cudaStream_t st) {
    size_t n = (size_t)c * hw;
    if (ac == c && ahw == hw && bc == c && (bhw == hw || bhw == 1) && hw > 1 && c <= 65535) {
        const int vec = hw % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0;
        const unsigned bx = min(cdiv(vec ? hw / 4 : hw, 256), (unsigned)max(1, (148 * 16) / c));
        if (bhw == 1) binary_rows_kernel<true><<<dim3(bx, c), 256, 0, st>>>(a, b, out, hw, op, vec);
        else binary_rows_kernel<false><<<dim3(bx, c), 256, 0, st>>>(a, b, out, hw, op, vec);
        g_launch_count++;
        return;
    }
    binary_kernel<<<min(cdiv(n, 256), 148u * 16), 256, 0, st>>>(a, ac, ahw, b, bc, bhw, out, c, hw, op);
    g_launch_count++;
}

// eltwise.cpp:79-150, op SUM with coefficients (two inputs): out = a*c0 + b*c1
__global__ void eltwise_sum2_kernel(const float* __restrict__ a, const float* __restrict__ b, float c0, float c1, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = a[i] * c0 + b[i] * c1;
}
void launch_eltwise_sum2(const float* a, const float* b, float c0, float c1, float* out, size_t n, cudaStream_t st) {
    eltwise_sum2_kernel<<<min(cdiv(n, 256), 148u * 16), 256, 0, st>>>(a, b, c0, c1, out, n);
    g_launch_count++;
}

// interp.cpp:54-91 (coefficients, computed in double then rounded to float) + :92-175 (H pass then V pass)
__device__ __forceinline__ void lin_coeff(int d, int in_n, int out_n, int& s, float& f) {
    double scale = (double)in_n / out_n;
    float fx = (float)((d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= in_n - 1) { sx = in_n - 2; fx = 1.f; }
    s = sx;
    f = fx;
}
__global__ void interp_bilinear_kernel(const float* __restrict__ in, int c, int h, int w, float* __restrict__ out, int oh, int ow) {
    int ox = blockIdx.x * blockDim.x + threadIdx.x;
    int oy = blockIdx.y;
    if (ox >= ow) return;
    int sx, sy;
    float fx, fy;
    lin_coeff(ox, w, ow, sx, fx);
    lin_coeff(oy, h, oh, sy, fy);
    float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
    for (int q = blockIdx.z; q < c; q += gridDim.z) {
        const float* p = in + (size_t)q * h * w;
        const float* r0 = p + (size_t)sy * w + sx;
        const float* r1 = r0 + w;
        float row0 = r0[0] * a0 + r0[1] * a1;
        float row1 = r1[0] * a0 + r1[1] * a1;
        out[((size_t)q * oh + oy) * ow + ox] = row0 * b0 + row1 * b1;
    }
}
void launch_interp_bilinear(const float* in, int c, int h, int w, float* out, int oh, int ow, cudaStream_t st) {
    dim3 grid(cdiv(ow, 128), oh, min(c, 16));
    interp_bilinear_kernel<<<grid, 128, 0, st>>>(in, c, h, w, out, oh, ow);
    g_launch_count++;
}

// pixelshuffle.cpp:33-80 mode 0: out[p][y*r+sh][x*r+sw] = in[p*r*r + sh*r + sw][y][x]
__global__ void pixelshuffle_kernel(const float* __restrict__ in, int c, int h, int w, float* __restrict__ out, int r) {
    int oc = c / (r * r), oh = h * r, ow = w * r;
    size_t n = (size_t)oc * oh * ow;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int ox = (int)(i % ow);
        size_t t = i / ow;
        int oy = (int)(t % oh), p = (int)(t / oh);
        int sh = oy % r, sw = ox % r;
        out[i] = in[((size_t)(p * r * r + sh * r + sw) * h + oy / r) * w + ox / r];
    }
}
void launch_pixelshuffle(const float* in, int c, int h, int w, float* out, int r, cudaStream_t st) {
    size_t n = (size_t)c * h * w;
    pixelshuffle_kernel<<<min(cdiv(n, 256), 148u * 16), 256, 0, st>>>(in, c, h, w, out, r);
    g_launch_count++;
}

// pooling.cpp:61-105 global average.  One block per (channel, slice): kPoolSlices partial sums per channel in a fixed order, then one
// warp per channel adds the slices (deterministic; a single block per channel left most of the machine idle on big planes).
constexpr int kPoolSlices = 16;
__device__ __forceinline__ float block_sum(float s, float* red) {
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    return s;
}
__global__ void global_avgpool_kernel(const float* __restrict__ in, float* __restrict__ out, size_t hw) {
    const float* p = in + (size_t)blockIdx.x * hw;
    float s = 0.f;
    for (size_t i = threadIdx.x; i < hw; i += blockDim.x) s += p[i];
    __shared__ float red[32];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[blockIdx.x] = s / (float)hw;
}
__global__ void avgpool_partial_kernel(const float* __restrict__ in, float* __restrict__ part, size_t hw) {
    const size_t per = (hw / 4 + kPoolSlices - 1) / kPoolSlices * 4;  // slice length, a multiple of 4 floats
    const size_t lo = (size_t)blockIdx.y * per, hi = min(hw, lo + per);
    const float* p = in + (size_t)blockIdx.x * hw;
    float s = 0.f;
    if ((hw & 3) == 0 && (((uintptr_t)in) & 15) == 0) {
        for (size_t i = lo / 4 + threadIdx.x; i < hi / 4; i += blockDim.x) { const float4 v = reinterpret_cast<const float4*>(p)[i]; s += (v.x + v.y) + (v.z + v.w); }
    } else {
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s += p[i];
    }
    __shared__ float red[32];
    s = block_sum(s, red);
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * kPoolSlices + blockIdx.y] = s;
}
__global__ void avgpool_final_kernel(const float* __restrict__ part, float* __restrict__ out, int c, size_t hw) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= c) return;
    float s = 0.f;
    for (int k = 0; k < kPoolSlices; k++) s += part[(size_t)q * kPoolSlices + k];
    out[q] = s / (float)hw;
}
// scratch: caller-owned room for c * kPoolSlices floats (one per executor, so concurrent lanes do not share it), or null
void launch_global_avgpool(const float* in, float* out, int c, size_t hw, cudaStream_t st, float* scratch) {
    if (scratch && hw >= 16384) {
        avgpool_partial_kernel<<<dim3(c, kPoolSlices), 256, 0, st>>>(in, scratch, hw);
        avgpool_final_kernel<<<cdiv(c, 128), 128, 0, st>>>(scratch, out, c, hw);
        g_launch_count += 2;
        return;
    }
    global_avgpool_kernel<<<c, 512, 0, st>>>(in, out, hw);
    g_launch_count++;
}
int global_avgpool_scratch_floats(int c) { return c * kPoolSlices; }

// innerproduct.cpp: out[p] = act(bias[p] + sum_i w[p][i]*x[i]) ; one warp per output
__global__ void innerproduct_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                                    int nin, int nout, int act, float p0) {
    int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= nout) return;
    float s = 0.f;
    for (int i = threadIdx.x & 31; i < nin; i += 32) s = fmaf(w[(size_t)p * nin + i], in[i], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) out[p] = apply_act(s + (bias ? bias[p] : 0.f), act, p0, 0.f);
}
void launch_innerproduct(const float* in, const float* w, const float* bias, float* out, int nin, int nout, int act, float p0, cudaStream_t st) {
    innerproduct_kernel<<<cdiv(nout, 4), 128, 0, st>>>(in, w, bias, out, nin, nout, act, p0);
    g_launch_count++;
}

// (the RIFE-specific HBM stages -- preproc, postproc, TTA averages, warp -- live in hbm_kernels.cu)
__global__ void fill_kernel(float* p, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
void launch_fill(float* p, size_t n, float v, cudaStream_t st) {
    fill_kernel<<<min(cdiv(n, 256), 148u * 16), 256, 0, st>>>(p, n, v);
    g_launch_count++;
}

}  // namespace rife
