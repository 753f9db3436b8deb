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
void launch_binary(const float* a, int ac, size_t ahw, const float* b, int bc, size_t bhw, float* out, int c, size_t hw, int op, cudaStream_t st) {
    size_t n = (size_t)c * hw;
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

// src/warp.cpp:96-168: backward bilinear warp, indices clamped, alpha/beta taken AFTER clamping.
__global__ void warp_kernel(const float* __restrict__ img, const float* __restrict__ flow, float* __restrict__ out, int c, int h, int w) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    size_t hw = (size_t)h * w, pi = (size_t)y * w + x;
    float sx = x + flow[pi], sy = y + flow[hw + pi];
    int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), w - 1);
    y0 = min(max(y0, 0), h - 1);
    x1 = min(max(x1, 0), w - 1);
    y1 = min(max(y1, 0), h - 1);
    float alpha = sx - x0, beta = sy - y0;
    for (int q = 0; q < c; q++) {
        const float* p = img + (size_t)q * hw;
        float v0 = p[(size_t)y0 * w + x0], v1 = p[(size_t)y0 * w + x1];
        float v2 = p[(size_t)y1 * w + x0], v3 = p[(size_t)y1 * w + x1];
        float v4 = v0 * (1 - alpha) + v1 * alpha;
        float v5 = v2 * (1 - alpha) + v3 * alpha;
        out[(size_t)q * hw + pi] = v4 * (1 - beta) + v5 * beta;
    }
}
void launch_warp(const float* img, const float* flow, float* out, int c, int h, int w, cudaStream_t st) {
    warp_kernel<<<dim3(cdiv(w, 128), h), 128, 0, st>>>(img, flow, out, c, h, w);
    g_launch_count++;
}

// pooling.cpp:61-105 global average
__global__ void global_avgpool_kernel(const float* __restrict__ in, float* __restrict__ out, size_t hw) {
    const float* p = in + (size_t)blockIdx.x * hw;
    float s = 0.f;
    for (size_t i = threadIdx.x; i < hw; i += blockDim.x) s += p[i];
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[blockIdx.x] = s / (float)hw;
    }
}
void launch_global_avgpool(const float* in, float* out, int c, size_t hw, cudaStream_t st) {
    global_avgpool_kernel<<<c, 512, 0, st>>>(in, out, hw);
    g_launch_count++;
}

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

// ------------------------------------------------------------------------------------------------
// RIFE stages
// ------------------------------------------------------------------------------------------------
// orientation maps (rife.cpp:3340-3364): destination index of padded source pixel (y,x)
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

// rife_preproc.comp:33-66 / rife.cpp:4152-4211: u8 -> float * (1/255), zero outside (w,h)
__global__ void preproc_kernel(const uint8_t* __restrict__ rgb, int w, int h, float* __restrict__ out, int wp, int hp, int orient) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= wp) return;
    float v[3] = {0.f, 0.f, 0.f};
    if (x < w && y < h) {
        const uint8_t* p = rgb + ((size_t)y * w + x) * 3;
        v[0] = (float)p[0] * (1 / 255.f);
        v[1] = (float)p[1] * (1 / 255.f);
        v[2] = (float)p[2] * (1 / 255.f);
    }
    size_t plane = (size_t)wp * hp, di = orient_index(orient, y, x, wp, hp);
    out[di] = v[0];
    out[plane + di] = v[1];
    out[2 * plane + di] = v[2];
}
void launch_preproc(const uint8_t* rgb, int w, int h, float* out, int wp, int hp, int orient, cudaStream_t st) {
    preproc_kernel<<<dim3(cdiv(wp, 128), hp), 128, 0, st>>>(rgb, w, h, out, wp, hp, orient);
    g_launch_count++;
}

__global__ void fill_kernel(float* p, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
void launch_fill(float* p, size_t n, float v, cudaStream_t st) {
    fill_kernel<<<min(cdiv(n, 256), 148u * 16), 256, 0, st>>>(p, n, v);
    g_launch_count++;
}

struct PostArgs {
    const float* in[16];
    int orient[16];
};
// rife_postproc.comp:33-63 / rife.cpp:4375-4398 + mat_pixel.cpp:158: v*255+0.5 -> (int) -> clamp -> u8.
// TTA: rife.cpp:4060-4144: mean of the 8 un-rotated outputs (/8), temporal: (v + vr) * 0.5.
__global__ void postproc_kernel(PostArgs pa, int n_in, int wp, int hp, uint8_t* __restrict__ rgb, int w, int h, int cpu_contig) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    size_t plane = (size_t)wp * hp;
    uint8_t* o = rgb + ((size_t)y * w + x) * 3;
    for (int q = 0; q < 3; q++) {
        float v;
        if (n_in == 1) {
            // the CPU reference reads the first w*h floats of each padded channel contiguously
            // (rife.cpp:4375-4387); identical to a proper crop whenever w == wp
            size_t idx = cpu_contig ? (size_t)y * w + x : (size_t)y * wp + x;
            v = pa.in[0][q * plane + idx] * 255.f + 0.5f;
        } else if (n_in == 2) {
            size_t idx = cpu_contig ? (size_t)y * w + x : (size_t)y * wp + x;
            v = (pa.in[0][q * plane + idx] + pa.in[1][q * plane + idx]) * 0.5f * 255.f + 0.5f;
        } else {
            float s = 0.f;
            for (int i = 0; i < 8; i++) s += pa.in[i][q * plane + orient_index(pa.orient[i], y, x, wp, hp)];
            s = s / 8;
            if (n_in == 16) {
                float sr = 0.f;
                for (int i = 8; i < 16; i++) sr += pa.in[i][q * plane + orient_index(pa.orient[i], y, x, wp, hp)];
                sr = sr / 8;
                v = (s + sr) * 0.5f * 255.f + 0.5f;
            } else {
                v = s * 255.f + 0.5f;
            }
        }
        int iv = (int)v;  // truncation, as mat_pixel.cpp:158 `(uchar)min(max((int)v,0),255)`
        o[q] = (uint8_t)min(max(iv, 0), 255);
    }
}
void launch_postproc(const float* const* ins, const int* orients, int n_in, int wp, int hp, uint8_t* rgb, int w, int h, int cpu_contig, cudaStream_t st) {
    PostArgs pa;
    for (int i = 0; i < 16; i++) {
        pa.in[i] = i < n_in ? ins[i] : nullptr;
        pa.orient[i] = i < n_in && orients ? orients[i] : 0;
    }
    postproc_kernel<<<dim3(cdiv(w, 128), h), 128, 0, st>>>(pa, n_in, wp, hp, rgb, w, h, cpu_contig);
    g_launch_count++;
}

// rife.cpp:2269-2319 (v1 rule) -- rife_flow_tta_temporal_avg.comp:19-42
__global__ void temporal_merge_v1_kernel(float* f, float* fr, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = (f[i] - fr[i]) * 0.5f, y = (f[n + i] - fr[n + i]) * 0.5f;
    f[i] = x; f[n + i] = y;
    fr[i] = -x; fr[n + i] = -y;
}
void launch_temporal_merge_v1(float* f, float* fr, size_t n, cudaStream_t st) {
    temporal_merge_v1_kernel<<<cdiv(n, 256), 256, 0, st>>>(f, fr, n);
    g_launch_count++;
}
// rife.cpp:2285-2306 (v2 rule), :4290-4311 (v4 adds the mask) -- rife_v2/v4_flow_tta_temporal_avg.comp
__global__ void temporal_merge_v2_kernel(float* f, float* fr, size_t n, int has_mask) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = (f[i] + fr[2 * n + i]) * 0.5f;
    float y = (f[n + i] + fr[3 * n + i]) * 0.5f;
    float z = (f[2 * n + i] + fr[i]) * 0.5f;
    float w = (f[3 * n + i] + fr[n + i]) * 0.5f;
    f[i] = x; f[n + i] = y; f[2 * n + i] = z; f[3 * n + i] = w;
    fr[i] = z; fr[n + i] = w; fr[2 * n + i] = x; fr[3 * n + i] = y;
    if (has_mask) {
        float m = (f[4 * n + i] - fr[4 * n + i]) * 0.5f;
        f[4 * n + i] = m;
        fr[4 * n + i] = -m;
    }
}
void launch_temporal_merge_v2(float* f, float* fr, size_t n, int has_mask, cudaStream_t st) {
    temporal_merge_v2_kernel<<<cdiv(n, 256), 256, 0, st>>>(f, fr, n, has_mask);
    g_launch_count++;
}

struct Flow8 {
    float* f[8];
};
// rife.cpp:1541-1719 (v1/v2), :3515-3668 (v4) -- rife_flow_tta_avg.comp / rife_v2_.. / rife_v4_..
// pair k of channels (2k, 2k+1) is an (x,y) flow; channel 4 (nch == 5) is the mask (plain mean).
__global__ void flow_tta_avg_kernel(Flow8 F, int nch, int fw, int fh) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;  // x in orientation 0
    int i = blockIdx.y;                             // y
    if (j >= fw) return;
    size_t plane = (size_t)fw * fh;
    size_t idx[8];
    for (int o = 0; o < 8; o++) idx[o] = orient_index(o, i, j, fw, fh);
    int npair = nch >= 4 ? 2 : 1;
    for (int k = 0; k < npair; k++) {
        size_t cx = (size_t)(2 * k) * plane, cy = (size_t)(2 * k + 1) * plane;
        float x = (F.f[0][cx + idx[0]] + -F.f[1][cx + idx[1]] + -F.f[2][cx + idx[2]] + F.f[3][cx + idx[3]] +
                   F.f[4][cy + idx[4]] + F.f[5][cy + idx[5]] + -F.f[6][cy + idx[6]] + -F.f[7][cy + idx[7]]) * 0.125f;
        float y = (F.f[0][cy + idx[0]] + F.f[1][cy + idx[1]] + -F.f[2][cy + idx[2]] + -F.f[3][cy + idx[3]] +
                   F.f[4][cx + idx[4]] + -F.f[5][cx + idx[5]] + -F.f[6][cx + idx[6]] + F.f[7][cx + idx[7]]) * 0.125f;
        F.f[0][cx + idx[0]] = x;  F.f[1][cx + idx[1]] = -x; F.f[2][cx + idx[2]] = -x; F.f[3][cx + idx[3]] = x;
        F.f[4][cx + idx[4]] = y;  F.f[5][cx + idx[5]] = -y; F.f[6][cx + idx[6]] = -y; F.f[7][cx + idx[7]] = y;
        F.f[0][cy + idx[0]] = y;  F.f[1][cy + idx[1]] = y;  F.f[2][cy + idx[2]] = -y; F.f[3][cy + idx[3]] = -y;
        F.f[4][cy + idx[4]] = x;  F.f[5][cy + idx[5]] = x;  F.f[6][cy + idx[6]] = -x; F.f[7][cy + idx[7]] = -x;
    }
    if (nch == 5) {
        size_t cm = 4 * plane;
        float m = 0.f;
        for (int o = 0; o < 8; o++) m += F.f[o][cm + idx[o]];
        m *= 0.125f;
        for (int o = 0; o < 8; o++) F.f[o][cm + idx[o]] = m;
    }
}
void launch_flow_tta_avg(float* const* f8, int nch, int fw, int fh, cudaStream_t st) {
    Flow8 F;
    for (int i = 0; i < 8; i++) F.f[i] = f8[i];
    flow_tta_avg_kernel<<<dim3(cdiv(fw, 128), fh), 128, 0, st>>>(F, nch, fw, fh);
    g_launch_count++;
}

}  // namespace rife
