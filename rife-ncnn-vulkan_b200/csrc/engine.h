// engine.h -- orchestration of one RIFE replica on one GPU: what RIFE::process_cpu / process_v4_cpu do in the
// reference (/root/reference/src/rife.cpp:1214-2460, 3204-4401), on device memory and CUDA streams.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "exec.h"
#include "model.h"

namespace rife {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    float* f() const { return (float*)p; }
    uint8_t* u8() const { return (uint8_t*)p; }
};

class Engine {
public:
    Engine(int gpuid, bool tta, bool tta_temporal, bool uhd, bool v2, bool v4);
    ~Engine();
    int init();  // selects the device, creates streams
    int load(const std::string& modeldir);
    int load_packed(const void* blob, size_t bytes);
    const std::string& packed() const { return packed_; }
    int process_host(const uint8_t* in0, const uint8_t* in1, int w, int h, float t, uint8_t* out);
    int process_device(const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out);
    int process_batch(int n, const uint8_t* const* in0, const uint8_t* const* in1, int w, int h, const float* ts, uint8_t* const* out);
    int set_option(const std::string& key, int value);
    void set_stream(cudaStream_t s) { std::lock_guard<std::mutex> lk(mu_); user_stream_ = s; use_user_stream_ = s != nullptr; }
    std::string last_error;

private:
    int finish_load();
    int run_device(const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out, cudaStream_t st);
    int run_v4(const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out, cudaStream_t st);
    int run_v1v2(const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, uint8_t* d_out, cudaStream_t st);
    Tensor keep(const Tensor& t, DevBuf& b, cudaStream_t st);  // copy a plan-owned tensor into an engine buffer

    int gpuid_;
    bool tta_, ttat_, uhd_, v2_, v4_;
    bool loaded_ = false;
    int precision_ = 1;
    bool async_ = false;
    cudaStream_t user_stream_ = nullptr;
    bool use_user_stream_ = false;
    Net nets_[3];           // flownet, contextnet, fusionnet
    NetRunner* run_[3] = {nullptr, nullptr, nullptr};
    std::string packed_;    // serialized model (param text + bin bytes per net)
    cudaStream_t st_ = nullptr, st_copy_[2] = {nullptr, nullptr};
    cudaEvent_t ev_[8] = {};
    std::mutex mu_;
    // device buffers
    DevBuf u8_[6];          // staged in0,in1,out (x2 for the pipelined batch path)
    void* pinned_[6] = {};  // pinned host staging
    size_t pinned_cap_[6] = {};
    DevBuf pad0_[8], pad1_[8], ts_[2], tsr_[2];
    DevBuf flow_[4][8], flowr_[4][8];
    DevBuf outp_[16];
    DevBuf ctx_[2][4];
    DevBuf tmp_[8];
};

}  // namespace rife
