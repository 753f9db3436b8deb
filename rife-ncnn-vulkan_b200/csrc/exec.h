// exec.h -- generic graph executor over the 22-op set the RIFE models use.
// Replaces ncnn::Net / ncnn::Extractor for this path: blobs are addressed by *name*, any blob (inputs or
// intermediates such as "flow0".."flow3") can be injected, and only the producers of missing blobs run --
// the behaviour the reference relies on in its TTA paths (src/ncnn/src/net.cpp:150-215, 2454-2500).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "model.h"

namespace rife {

struct Tensor {
    float* p = nullptr;
    int dims = 0;  // 1: (w) vector, 3: (c,h,w)
    int c = 0, h = 0, w = 0;
    size_t count() const { return dims == 1 ? (size_t)w : (size_t)c * h * w; }
    static Tensor chw(float* p, int c, int h, int w) {
        Tensor t;
        t.p = p; t.dims = 3; t.c = c; t.h = h; t.w = w;
        return t;
    }
};

struct DeviceWeights {
    float* wT = nullptr;    // conv: [Cin][KK][ocpad]; deconv: [4][Cin][4][ocpad]; innerproduct: [nout][nin]
    float* bias = nullptr;
    float* slope = nullptr;
    int ocpad = 0;
    // tensor-core path (tc_conv.cu): packed fp16 weights, bias padded to the GEMM N
    void* wpk = nullptr;
    float* biasN = nullptr;
    int tcN = 0, ocs = 0, cin = 0, cinp = 0, tc_s2 = 0, tc_k5 = 0;
    int nchunks = 1;         // wide layers run as nchunks launches of tcN GEMM columns each (output-channel slices)
    size_t chunk_elems = 0;  // packed fp16 elements per chunk
};

class NetRunner {
public:
    NetRunner() {}
    ~NetRunner();
    int init(const Net* net, std::string& err);  // uploads weights to the current device
    // second and further runners of the same net on the same device (concurrent lanes) borrow the owner's weights
    void share_from(const NetRunner& owner) { net_ = owner.net_; dwp_ = owner.dwp_; owns_ = false; tc_mode = owner.tc_mode; num_sms = owner.num_sms; fuse = owner.fuse; }
    const Net* net() const { return net_; }
    bool fuse = true;  // conv+add+leaky / conv+prelu epilogue fusion
    int tc_mode = 1;   // 0: fp32 CUDA-core kernels only; 1: tcgen05 with split-fp16 (hi+lo) activations; 2: tcgen05, plain fp16
    int num_sms = 148;
    void clear_plans();

    // Runs the sub-graph needed for `outputs` given `inputs`.  Output tensors point into plan-owned memory that
    // stays valid until the next run() of the same (inputs-shape, outputs) signature.
    int run(const std::vector<std::pair<std::string, Tensor>>& inputs, const std::vector<std::string>& outputs,
            std::vector<Tensor>& out_tensors, cudaStream_t st, std::string& err);

    size_t arena_bytes() const;
    const DeviceWeights& weights(int layer) const { return (*dwp_)[layer]; }

private:
    struct Step {
        int kind = 0;              // 0 generic layer, 1 tcgen05 conv / deconv, 2 planar -> C8, 3 C8 -> planar
        int layer = -1;
        int fused_add_blob = -1;   // residual blob id fused into the conv epilogue
        int fused_act_layer = -1;  // ReLU / PReLU layer index fused after
        int fused_ps_layer = -1;   // PixelShuffle fused into the tensor-core deconv epilogue
        int out_blob = -1;         // blob written (differs from layer top when fused)
        int conv_root = -1;        // kinds 2/3: root blob converted
    };
    struct Plan {
        std::vector<Step> steps;
        std::vector<Tensor> blobs;       // shape + (arena-relative) pointer per blob id
        std::vector<size_t> offset;      // arena offset (planar fp32 storage) or (size_t)-1
        std::vector<size_t> offset_c8;   // arena offset of the C8 fp16 storage of a root blob or (size_t)-1
        int split = 0;                   // C8 tensors carry a lo plane
        std::vector<char> c8_s2d;        // root blob's C8 storage is in space-to-depth form (feeds a stride-2 tensor-core conv)
        std::vector<int> external_slot;  // blob id -> index into inputs, or -1
        float* arena = nullptr;
        size_t arena_size = 0;
        std::vector<int> out_ids;
        std::vector<int> root;    // storage root of each blob (aliases: Split tops, channel Crops)
        std::vector<size_t> eoff; // element offset of the blob inside its root
    };
    int build_plan(const std::vector<std::pair<std::string, Tensor>>& inputs, const std::vector<std::string>& outputs, Plan& plan, std::string& err);
    int exec_step(Plan& plan, const Step& s, cudaStream_t st, std::string& err);

    const Net* net_ = nullptr;
    std::vector<DeviceWeights> dw_own_;
    const std::vector<DeviceWeights>* dwp_ = nullptr;
    bool owns_ = true;
    std::map<std::string, std::unique_ptr<Plan>> plans_;
    float* pool_scratch_ = nullptr;  // partial sums of the global average pools (per executor: lanes run concurrently)
    int pool_scratch_c_ = 0;
};

}  // namespace rife
