// kernels.h -- host launchers for the fp32 planar (CHW) kernels of the generic executor and for the
// RIFE-specific HBM kernels (preproc / postproc / warp / TTA averages).  All launch on the given stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rife {

// counts kernel launches issued by this library (reported by bench.py as gpu_launches)
extern unsigned long long g_launch_count;
extern unsigned long long g_h2d_bytes, g_d2h_bytes;

struct ConvArgs {
    const float* in;
    const float* wT;    // [parity][Cin][KK][ocpad]
    const float* bias;  // may be null
    float* out;
    int Cin, H, W;
    int Cout, OH, OW;
    int DH, DW;                // compute domain
    int in_off_y, in_off_x;    // input coord = d*S + off + k
    int out_mul, out_off_y, out_off_x;
    int act;                   // ncnn activation_type (0 none, 1 relu, 2 leaky, 3 clip, 4 sigmoid)
    float act_p0, act_p1;
    int ocpad;                 // padded Cout of wT (multiple of 64)
    int nparity;               // 1 for conv; 4 for deconv4x4s2 (parity p -> out_off = (p>>1, p&1), in_off = off + (p>>1, p&1))
    // optional fused epilogue:  v = act(conv + bias); if (res) v += res[...]; if (post_act) v = leaky/prelu(v)
    const float* res;          // residual, same shape as out (or null)
    int post_act;              // 0 none, 2 leaky with post_p0, 5 per-channel prelu with post_slope
    float post_p0;
    const float* post_slope;
};
void launch_conv(const ConvArgs& a, int K, int S, cudaStream_t st);

enum UnaryOp { U_RELU = 0, U_LEAKY, U_SIGMOID, U_CLIP, U_NEG, U_ADD_S, U_SUB_S, U_MUL_S, U_DIV_S, U_RSUB_S, U_RDIV_S, U_COPY };
void launch_unary(const float* in, float* out, size_t n, int op, float p0, float p1, cudaStream_t st);
void launch_prelu(const float* in, const float* slope, int nslope, float* out, int c, size_t hw, cudaStream_t st);

enum BinOp { B_ADD = 0, B_SUB, B_MUL, B_DIV, B_MAX, B_MIN, B_POW, B_RSUB, B_RDIV };
// a: (ac, ahw) b: (bc, bhw); each of ac/bc is 1 or c, each of ahw/bhw is 1 or hw
void launch_binary(const float* a, int ac, size_t ahw, const float* b, int bc, size_t bhw, float* out, int c, size_t hw, int op, cudaStream_t st);
void launch_eltwise_sum2(const float* a, const float* b, float c0, float c1, float* out, size_t n, cudaStream_t st);
void launch_interp_bilinear(const float* in, int c, int h, int w, float* out, int oh, int ow, cudaStream_t st);
void launch_pixelshuffle(const float* in, int c, int h, int w, float* out, int r, cudaStream_t st);
void launch_warp(const float* img, const float* flow, float* out, int c, int h, int w, cudaStream_t st);
void launch_global_avgpool(const float* in, float* out, int c, size_t hw, cudaStream_t st, float* scratch = nullptr);
int global_avgpool_scratch_floats(int c);
void launch_innerproduct(const float* in, const float* w, const float* bias, float* out, int nin, int nout, int act, float p0, cudaStream_t st);

// ---- RIFE stages, hbm_kernels.cu ----
// u8 HWC -> planar float * (1/255), zero outside (w,h), for the first `norient` (1 or 8) TTA orientations (Appendix B index
// maps) from one read of the frame: orientation o at out + o * 3 * wp * hp, as [3][hp][wp] (o < 4) or [3][wp][hp] (o >= 4).
// bgr: the frame bytes are B,G,R (the reference's Windows build); the planes are always R,G,B.
void launch_preproc(const uint8_t* rgb, int w, int h, float* out, int wp, int hp, int norient, int bgr, cudaStream_t st);
void launch_fill(float* p, size_t n, float v, cudaStream_t st);
// planar float -> u8 HWC: v*255+0.5, trunc, clamp.  n_in = 1 (plain), 2 (temporal TTA: mean of two), 8 / 16 (spatial TTA:
// input i is orientation i & 7, un-rotated and averaged; 16 = both temporal directions).
// cpu_contig (n_in <= 2 only): read the first w*h floats of every padded channel contiguously, as the reference's CPU path
// does (rife.cpp:4375-4387), instead of cropping the padded rows.
void launch_postproc(const float* const* ins, int n_in, int wp, int hp, uint8_t* rgb, int w, int h, int cpu_contig, int bgr, cudaStream_t st);
// temporal merges (in place), n = elements per channel
void launch_temporal_merge_v1(float* f, float* fr, size_t n, cudaStream_t st);            // 2 ch: (x - xr)/2 ; rev = -x
void launch_temporal_merge_v2(float* f, float* fr, size_t n, int has_mask, cudaStream_t st); // 4 ch (+ mask at ch 4)
// spatial TTA flow average over 8 orientation blobs (in place). nch = 2 (v1), 4 (v2), 5 (v4: 4 flow + mask)
// blobs 0-3 are [nch_total][fh][fw], blobs 4-7 are [nch_total][fw][fh].
void launch_flow_tta_avg(float* const* f8, int nch, int fw, int fh, cudaStream_t st);

}  // namespace rife
