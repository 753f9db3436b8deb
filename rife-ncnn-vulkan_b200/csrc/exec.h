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
};

class NetRunner {
public:
    NetRunner() {}
    ~NetRunner();
    int init(const Net* net, std::string& err);  // uploads weights to the current device
    const Net* net() const { return net_; }
    bool fuse = true;  // conv+add+leaky / conv+prelu epilogue fusion

    // Runs the sub-graph needed for `outputs` given `inputs`.  Output tensors point into plan-owned memory that
    // stays valid until the next run() of the same (inputs-shape, outputs) signature.
    int run(const std::vector<std::pair<std::string, Tensor>>& inputs, const std::vector<std::string>& outputs,
            std::vector<Tensor>& out_tensors, cudaStream_t st, std::string& err);

    size_t arena_bytes() const;

private:
    struct Step {
        int layer;
        int fused_add_blob = -1;   // residual blob id fused into the conv epilogue
        int fused_act_layer = -1;  // ReLU / PReLU layer index fused after
        int out_blob = -1;         // blob written (differs from layer top when fused)
    };
    struct Plan {
        std::vector<Step> steps;
        std::vector<Tensor> blobs;       // shape + (arena-relative) pointer per blob id
        std::vector<size_t> offset;      // arena offset or (size_t)-1 for external
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
    std::vector<DeviceWeights> dw_;
    std::map<std::string, std::unique_ptr<Plan>> plans_;
};

}  // namespace rife
