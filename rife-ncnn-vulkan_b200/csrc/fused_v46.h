// fused_v46.h -- hand-scheduled fast path for the rife-v4.6 IFNet (see fused_v46.cu)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "exec.h"
#include "model.h"

namespace rife {

constexpr int V46_MAX_BATCH = 8;

class V46Runner {
public:
    ~V46Runner();
    // `weights` = the NetRunner that owns the packed tensor-core weights of `net` on this device
    int init(const Net* net, const NetRunner* weights, std::string& err);
    bool ok() const { return ok_; }
    // bit k set: the residual chain of IFBlock k runs on plain fp16 activations (half the tensor work, ~2^-11 operand
    // rounding); 0 = every tensor-core activation is split hi+lo
    void set_plain_mask(int m) { plain_mask_ = m & 255; }
    // 0: flow / mask planes stored at full resolution between the block-2 head and the tail; 1: block-3 head stops
    // storing them; 2: never stored, rebuilt from the per-block flow tensors where needed (same arithmetic, fewer HBM bytes)
    void set_recompute(int m) { recompute_ = m < 0 ? 0 : (m > 2 ? 2 : m); }
    int run(const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out, cudaStream_t st, std::string& err);
    int run_batch(int n, const uint8_t* const* d_in0, const uint8_t* const* d_in1, int w, int h, const float* ts, uint8_t* const* d_out, cudaStream_t st,
                  std::string& err);

private:
    int ensure(int w, int h, int batch, std::string& err);
    int conv(int layer, const __half* in, __half* out, const __half* res, float* out_f32, int oh, int ow, bool out_s2d, int batch, bool split_in,
             bool split_out, cudaStream_t st);
    const Net* net_ = nullptr;
    const NetRunner* wr_ = nullptr;
    bool ok_ = false;
    float slope_ = 0.2f;
    int plain_mask_ = 0;
    int recompute_ = 0;
    std::vector<int> conv_;  // the 44 conv / deconv layer indices in graph order
    int wp_ = 0, hp_ = 0, cap_ = 0;
    std::vector<void*> bufs_;
    uchar4* rgbx_ = nullptr;  // the distinct frames of a batch, padded RGBX
    float *F_ = nullptr, *M_ = nullptr, *d_[4] = {};
    __half *x_[4] = {}, *y0_[4] = {}, *a_[4] = {}, *b_[4] = {};
};

}  // namespace rife
