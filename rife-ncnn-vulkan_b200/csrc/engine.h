// engine.h -- orchestration of one RIFE replica on one GPU: what RIFE::process_cpu / process_v4_cpu do in the
// reference (/root/reference/src/rife.cpp:1214-2460, 3204-4401), on device memory and CUDA streams.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "combiner.h"
#include "exec.h"
#include "fused_v46.h"
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

// one concurrent execution context: own stream, own plans (arenas) and scratch; weights are shared
struct Lane {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t fork = nullptr;  // spatial TTA: the coordinator's fork point for its helper lanes
    NetRunner* run[3] = {nullptr, nullptr, nullptr};
    V46Runner* fast = nullptr;  // hand-scheduled rife-v4 / v4.6 path (plain mode, precision tier 1)
    DevBuf pad0, pad1;          // the 8 orientations of the two padded frames, planar fp32 (generic path)
    DevBuf ts[2], tsr[2];
    DevBuf flow[4][8], flowr[4][8];
    DevBuf outp[16];
    DevBuf ctx[2][4];
    DevBuf tmp[8];
    void release();
};

constexpr int RECOMPUTE_FM_DEFAULT = 0;
constexpr int HEAD_PACK_DEFAULT = 0;
constexpr int COMBINE_DEFAULT = 1;  // measured (profiles/r1_s21/process_threads_1080p.txt): 8 caller threads 668 -> 1091 process() calls/s at 1080p

class Engine {
public:
    Engine(int gpuid, bool tta, bool tta_temporal, bool uhd, bool v2, bool v4);
    ~Engine();
    int init();  // selects the device, creates streams
    int load(const std::string& modeldir);
    int load_packed(const void* blob, size_t bytes);
    std::string packed() const { std::lock_guard<std::mutex> lk(mu_); return packed_; }
    int process_host(const uint8_t* in0, const uint8_t* in1, int w, int h, float t, uint8_t* out);
    int process_device(const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out);
    int process_batch(int n, const uint8_t* const* in0, const uint8_t* const* in1, int w, int h, const float* ts, uint8_t* const* out);
    int process_batch_device(int n, const uint8_t* const* d_in0, const uint8_t* const* d_in1, int w, int h, const float* ts, uint8_t* const* d_out);
    int set_option(const std::string& key, int value);
    struct HostReq {  // one process() call waiting in the combiner
        const uint8_t* in0;
        const uint8_t* in1;
        int w, h;
        float t;
        uint8_t* out;
    };
    int get_option(const std::string& key, int* value);
    void set_stream(cudaStream_t s) { std::lock_guard<std::mutex> lk(mu_); user_stream_ = s; use_user_stream_ = s != nullptr; }
    std::string stage_report();         // option "ktime": per-lane, per-stage times of the fused path
    void forget_frames();               // drops every cached input frame (option "frame_cache")
    std::string last_error() const { std::lock_guard<std::mutex> lk(err_mu_); return last_error_; }

private:
    int run_device(Lane& L, const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out, cudaStream_t st);
    int run_v4(Lane& L, const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out, cudaStream_t st);
    int run_v1v2(Lane& L, const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, uint8_t* d_out, cudaStream_t st);
    int make_lanes(int n);
    void tta_fork(Lane& L);
    void tta_join(Lane& L);
    Lane& tta_lane(Lane& L, int job);
    void setup_fast();      // (re)creates the per-lane fast runners and validates them against the generic executor
    Tensor keep(const Tensor& t, DevBuf& b, cudaStream_t st);  // copy a plan-owned tensor into an engine buffer
    void set_error(const std::string& s) { std::lock_guard<std::mutex> lk(err_mu_); last_error_ = s; }
    void sync_all();        // error paths: nothing of this handle may still touch caller buffers after an error return
    void publish();         // refreshes the lock-free snapshot process_host() reads (call with mu_ held)
    bool fast_usable() const { return fast_ok_ && use_fast_ && v4_ && !tta_ && !ttat_ && precision_ == 1; }

    int gpuid_;
    bool tta_, ttat_, uhd_, v2_, v4_;
    bool loaded_ = false;
    int precision_ = 1;
    bool async_ = false;
    bool fast_ok_ = false;  // the fused path reproduced the generic executor on the self-check
    int use_fast_ = 1;
    int plain_mask_ = 12;  // IFBlocks 2 and 3 (80 % of the FLOPs): plain fp16 activations in the residual chain (profiles/r1_precision_study_plain_blocks.txt)
    int combine_ = COMBINE_DEFAULT;  // concurrent process() calls on this handle are executed as one lock-step batch (combiner.h)
    int cpu_crop_quirk_ = 0;  // 1: reproduce the reference CPU path's contiguous read of the padded output (rife.cpp:4375-4387)
    int head_pack_ = HEAD_PACK_DEFAULT;  // fused path: packed block-head tensors (fused_v46_kernels.cuh)
    int d2h_on_lane_ = 0;     // process_batch: 1 = a chunk's results go home on the lane's own stream instead of the two copy streams
                              // (RIFE_B200_D2H; measured 2 % slower at 1080p, equal at 4K: profiles/README.md, session 19)
    int bgr_ = 0;             // frames are B,G,R in memory (the reference's Windows build: rife_preproc.comp:13,53-56)
    Combiner<HostReq> combiner_;
    int run_combined(HostReq** rq, int n);
    int recompute_fm_ = RECOMPUTE_FM_DEFAULT;  // fused path: rebuild the full-resolution flow / mask planes instead of storing them (0, 1, 2: fused_v46.h)
    cudaStream_t user_stream_ = nullptr;
    bool use_user_stream_ = false;
    std::unique_ptr<Net> nets_[3];  // flownet, contextnet, fusionnet
    std::vector<Lane*> lanes_;
    int nlanes_ = 2;
    std::string packed_;    // serialized model (param text + bin bytes per net)
    cudaStream_t st_copy_[3] = {nullptr, nullptr, nullptr};  // [0] H2D, [1] / [2] D2H (chunks alternate: two copies in flight keep the link busier than one engine's queue)
    static const int kSlots = 8;
    cudaEvent_t ev_h2d_[kSlots] = {}, ev_comp_[kSlots] = {}, ev_d2h_[kSlots] = {}, ev_entry_ = nullptr;
    mutable std::mutex mu_;
    mutable std::mutex err_mu_;
    std::string last_error_;
    // What process_host() needs to decide whether a call goes through the combiner, readable without mu_ (the leader of a
    // combined batch holds mu_ for the whole batch; followers must be able to queue meanwhile).
    std::atomic<int> snap_combine_{0}, snap_fast_{0}, snap_batch_{0}, snap_lanes_{1};
    // device buffers
    std::vector<DevBuf> out_u8_;  // staged outputs per pipeline slot and batch position
    // Input frames on the device, found again by host pointer.  Within one call a frame shared by
    // several pairs is uploaded once; with option "frame_cache" = 1 entries survive across calls (the caller promises not to
    // modify a frame buffer it has handed in until it calls forget_frames / turns the option off).
    struct FrameEntry {
        const uint8_t* host = nullptr;
        size_t nb = 0;
        DevBuf buf;
        cudaEvent_t read_done = nullptr;  // last compute that read buf (recorded on a lane stream)
        bool reading = false;
        unsigned long long stamp = 0;
    };
    std::vector<FrameEntry> frames_;
    unsigned long long frame_clock_ = 0;
    int frame_cache_ = 0;
    unsigned long long frame_hits_ = 0;
    FrameEntry* frame_lookup(const uint8_t* host, size_t nb, FrameEntry* const* cur, int ncur, bool* hit);
    void frames_fit(size_t nb);
    // Pinned staging for process() calls that arrive with pageable buffers (what the reference CLI passes: stb_image / malloc
    // memory, src/main.cpp:140-187): every caller thread copies its own frames into a pinned slot BEFORE it queues, and its result
    // out of the slot afterwards, so the copies of concurrent callers run in parallel on their own cores and the thread that
    // executes the combined batch only issues asynchronous DMA (a pageable cudaMemcpyAsync is staged by the driver, serially,
    // in that one thread).  Option "stage_pageable" (default 1; ignored while "frame_cache" is on: slots are reused, pointers
    // would lie).
    struct StageSlot { uint8_t* p = nullptr; size_t cap = 0; bool busy = false; };
    std::vector<StageSlot> stage_;
    std::mutex stage_mu_;
    int stage_pageable_ = 1;
    std::atomic<int> snap_stage_{1};
    uint8_t* stage_acquire(size_t bytes, int* idx);
    void stage_release(int idx);
    int batch_ = 0;           // pairs per lock-step batch on the fused path (0 = choose from the frame size)
    int batch_for(int w, int h) const;
    int run_chunk(Lane& L, int n, const uint8_t* const* d_in0, const uint8_t* const* d_in1, int w, int h, const float* ts, uint8_t* const* d_out, cudaStream_t st);
};

}  // namespace rife
