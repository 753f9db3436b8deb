// engine.cu -- see engine.h.  Dataflow restated from /root/reference/src/rife.cpp (line refs per block).
#include "engine.h"

#include <stdio.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "kernels.h"

namespace rife {

int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return -1;
    cap = bytes;
    return 0;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

Engine::Engine(int gpuid, bool tta, bool tta_temporal, bool uhd, bool v2, bool v4)
    : gpuid_(gpuid), tta_(tta), ttat_(tta_temporal), uhd_(uhd), v2_(v2), v4_(v4) {
    // process-wide default of option "recompute_fm" (lets a whole test run exercise one setting)
    if (const char* e = getenv("RIFE_B200_RECOMPUTE_FM")) { int v = atoi(e); recompute_fm_ = v < 0 ? 0 : (v > 2 ? 2 : v); }
    if (const char* e = getenv("RIFE_B200_COMBINE")) combine_ = atoi(e) != 0;
    if (const char* e = getenv("RIFE_B200_D2H")) d2h_on_lane_ = atoi(e) != 0;
    if (tta_) nlanes_ = 8;  // the 8 orientations of a pair are dealt to the lanes (tta_fork / tta_join): measured 7.1 (1 lane) / 13.8 (4) / 14.7 (8) fps for rife-anime 1080p -x -z
}

void Lane::release() {
    delete fast;
    fast = nullptr;
    for (auto& r : run) { delete r; r = nullptr; }
    pad0.release();
    pad1.release();
    for (int i = 0; i < 8; i++) tmp[i].release();
    for (int i = 0; i < 2; i++) { ts[i].release(); tsr[i].release(); for (auto& c : ctx[i]) c.release(); }
    for (auto& a : flow) for (auto& b : a) b.release();
    for (auto& a : flowr) for (auto& b : a) b.release();
    for (auto& b : outp) b.release();
    if (done) cudaEventDestroy(done);
    if (fork) cudaEventDestroy(fork);
    fork = nullptr;
    if (st) cudaStreamDestroy(st);
    done = nullptr;
    st = nullptr;
}

Engine::~Engine() {
    cudaSetDevice(gpuid_);
    cudaDeviceSynchronize();
    // borrowers first, the weight-owning lane 0 last
    for (size_t i = lanes_.size(); i-- > 0;) { lanes_[i]->release(); delete lanes_[i]; }
    for (auto& b : out_u8_) b.release();
    for (auto& sl : stage_) if (sl.p) cudaFreeHost(sl.p);
    for (auto& f : frames_) { f.buf.release(); if (f.read_done) cudaEventDestroy(f.read_done); }
    for (int i = 0; i < kSlots; i++) {
        if (ev_h2d_[i]) cudaEventDestroy(ev_h2d_[i]);
        if (ev_comp_[i]) cudaEventDestroy(ev_comp_[i]);
        if (ev_d2h_[i]) cudaEventDestroy(ev_d2h_[i]);
    }
    if (ev_entry_) cudaEventDestroy(ev_entry_);
    for (auto& s : st_copy_) if (s) cudaStreamDestroy(s);
}

int Engine::init() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || gpuid_ < 0 || gpuid_ >= n) { set_error("no such CUDA device"); return -2; }
    if (cudaSetDevice(gpuid_) != cudaSuccess) { set_error("cudaSetDevice failed"); return -2; }
    for (auto& s : st_copy_) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int i = 0; i < kSlots; i++) {
        cudaEventCreateWithFlags(&ev_h2d_[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_comp_[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_d2h_[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ev_entry_, cudaEventDisableTiming);
    return make_lanes(1);
}

// (re)creates the lanes; lane 0 keeps (owns) the weights, the others borrow them
int Engine::make_lanes(int n) {
    if (n < 1) n = 1;
    if (n > 8) n = 8;
    while ((int)lanes_.size() > n) { lanes_.back()->release(); delete lanes_.back(); lanes_.pop_back(); }
    while ((int)lanes_.size() < n) {
        Lane* L = new Lane();
        if (cudaStreamCreateWithFlags(&L->st, cudaStreamNonBlocking) != cudaSuccess) { delete L; set_error("stream creation failed"); return -2; }
        cudaEventCreateWithFlags(&L->done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&L->fork, cudaEventDisableTiming);
        if (!lanes_.empty())
            for (int i = 0; i < 3; i++)
                if (lanes_[0]->run[i]) { L->run[i] = new NetRunner(); L->run[i]->share_from(*lanes_[0]->run[i]); }
        lanes_.push_back(L);
    }
    return 0;
}

static const char* kNetNames[3] = {"flownet", "contextnet", "fusionnet"};

static bool read_file(const std::string& path, std::string& out) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    out.resize((size_t)n);
    size_t rd = n ? fread(&out[0], 1, (size_t)n, fp) : 0;
    fclose(fp);
    return rd == (size_t)n;
}

// packed model = "RIFEB200" u32 nnets { u64 param_len, u64 bin_len, param bytes, bin bytes }*
int Engine::load(const std::string& modeldir) {
    int nn = v4_ ? 1 : 3;  // rife.cpp:158-163
    std::string blob = "RIFEB200";
    uint32_t n32 = (uint32_t)nn;
    blob.append((const char*)&n32, 4);
    for (int i = 0; i < nn; i++) {
        std::string p, b;
        if (!read_file(modeldir + "/" + kNetNames[i] + ".param", p) || !read_file(modeldir + "/" + kNetNames[i] + ".bin", b)) {
            set_error("cannot read " + modeldir + "/" + kNetNames[i] + ".{param,bin}");
            return -3;
        }
        uint64_t pl = p.size(), bl = b.size();
        blob.append((const char*)&pl, 8);
        blob.append((const char*)&bl, 8);
        blob += p;
        blob += b;
    }
    return load_packed(blob.data(), blob.size());
}

// Transactional: the new model is parsed and uploaded into temporaries; the engine's state changes only after every
// network loaded, so a failing (re)load leaves a previously loaded engine exactly as it was.
int Engine::load_packed(const void* data, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu_);
    cudaSetDevice(gpuid_);
    const char* p = (const char*)data;
    if (bytes < 12 || memcmp(p, "RIFEB200", 8)) { set_error("bad packed model"); return -3; }
    uint32_t nn;
    memcpy(&nn, p + 8, 4);
    if (nn != (uint32_t)(v4_ ? 1 : 3)) { set_error("packed model does not match the model family flags"); return -3; }
    size_t pos = 12;
    std::unique_ptr<Net> nets[3];
    std::unique_ptr<NetRunner> runs[3];
    int num_sms = 148;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, gpuid_) == cudaSuccess) num_sms = prop.multiProcessorCount;
    }
    for (uint32_t i = 0; i < nn; i++) {
        uint64_t pl, bl;
        if (bytes - pos < 16) { set_error("truncated packed model"); return -3; }
        memcpy(&pl, p + pos, 8);
        memcpy(&bl, p + pos + 8, 8);
        pos += 16;
        if (pl > bytes - pos || bl > bytes - pos - pl) { set_error("truncated packed model"); return -3; }  // no u64 wrap-around
        std::string ptxt(p + pos, (size_t)pl), bbin(p + pos + pl, (size_t)bl);
        pos += (size_t)pl + (size_t)bl;
        std::string err;
        nets[i].reset(new Net());
        if (parse_net(ptxt, bbin, kNetNames[i], *nets[i], err)) { set_error(err); return -3; }
        runs[i].reset(new NetRunner());
        runs[i]->tc_mode = precision_;
        runs[i]->num_sms = num_sms;
        if (runs[i]->init(nets[i].get(), err)) { set_error(err); cudaGetLastError(); return -3; }
    }
    // commit
    cudaDeviceSynchronize();
    loaded_ = false;
    fast_ok_ = false;
    publish();
    for (Lane* L : lanes_) { delete L->fast; L->fast = nullptr; }
    for (size_t li = lanes_.size(); li-- > 0;)  // borrowers first
        for (int i = 0; i < 3; i++) { delete lanes_[li]->run[i]; lanes_[li]->run[i] = nullptr; }
    for (uint32_t i = 0; i < 3; i++) {
        nets_[i] = std::move(nets[i]);
        if (i >= nn) continue;
        lanes_[0]->run[i] = runs[i].release();
        for (size_t li = 1; li < lanes_.size(); li++) { lanes_[li]->run[i] = new NetRunner(); lanes_[li]->run[i]->share_from(*lanes_[0]->run[i]); }
    }
    packed_.assign(p, bytes);
    loaded_ = true;
    if ((int)lanes_.size() != nlanes_) { int r = make_lanes(nlanes_); if (r) { loaded_ = false; publish(); return r; } }
    setup_fast();
    publish();
    return 0;
}

void Engine::publish() {
    snap_combine_.store(combine_ && loaded_ ? 1 : 0, std::memory_order_relaxed);
    snap_fast_.store(loaded_ && fast_usable() ? 1 : 0, std::memory_order_relaxed);
    snap_batch_.store(batch_, std::memory_order_relaxed);
    snap_stage_.store(stage_pageable_ && !frame_cache_ ? 1 : 0, std::memory_order_relaxed);
    snap_lanes_.store((int)lanes_.size(), std::memory_order_release);
}

void Engine::sync_all() {
    for (auto& s : st_copy_) if (s) cudaStreamSynchronize(s);
    for (Lane* L : lanes_) if (L->st) cudaStreamSynchronize(L->st);
    if (use_user_stream_) cudaStreamSynchronize(user_stream_);
    cudaGetLastError();
}

int Engine::set_option(const std::string& key, int value) {
    std::lock_guard<std::mutex> lk(mu_);
    struct Pub { Engine* e; ~Pub() { e->publish(); } } pub{this};
    if (key == "precision") {  // 0 exact fp32, 1 tensor cores + split-fp16 activations, 2 tensor cores + plain fp16 activations
        precision_ = value;
        cudaDeviceSynchronize();
        for (Lane* L : lanes_) for (auto& r : L->run) if (r) { r->tc_mode = value; r->clear_plans(); }
        return 0;
    }
    if (key == "lanes") {
        nlanes_ = value;
        cudaDeviceSynchronize();
        if (!loaded_) return 0;
        int r = make_lanes(value);
        if (r) return r;
        setup_fast();
        return 0;
    }
    if (key == "fast") { use_fast_ = value; return 0; }
    if (key == "plain_blocks") {  // bit mask of IFBlocks whose residual chain uses plain fp16 activations (fused path)
        plain_mask_ = value & 255;  // bits 4-7 (experimental): block k's head tensor / conv0 input in plain fp16
        for (Lane* L : lanes_) if (L->fast) L->fast->set_plain_mask(plain_mask_);
        return 0;
    }
    if (key == "recompute_fm") {  // fused path: 0 store the full-resolution flow / mask planes, 1 / 2 rebuild them from the block outputs (fused_v46.h)
        recompute_fm_ = value < 0 ? 0 : (value > 2 ? 2 : value);
        for (Lane* L : lanes_) if (L->fast) L->fast->set_recompute(recompute_fm_);
        return 0;
    }
    if (key == "cpu_crop_quirk") {  // 1: the output is the reference CPU path's contiguous read of the padded planes (rife.cpp:4375-4387)
        cpu_crop_quirk_ = value != 0;
        for (Lane* L : lanes_) if (L->fast) L->fast->set_crop_quirk(cpu_crop_quirk_);
        return 0;
    }
    if (key == "bgr") {  // frames are B,G,R in memory (the reference's Windows build)
        bgr_ = value != 0;
        for (Lane* L : lanes_) if (L->fast) L->fast->set_bgr(bgr_);
        for (auto& f : frames_) f.host = nullptr;
        return 0;
    }
    if (key == "frame_cache") {  // 1: uploaded input frames stay on the device across calls, found again by host pointer
        frame_cache_ = value != 0;
        if (!frame_cache_) for (auto& f : frames_) f.host = nullptr;
        return 0;
    }
    if (key == "head_pack") {  // fused path: packed block-head tensors (32 instead of 64 bytes per pixel; fused_v46_kernels.cuh)
        head_pack_ = value != 0;
        for (Lane* L : lanes_) if (L->fast) L->fast->set_head_pack(head_pack_);
        return 0;
    }
    if (key == "ktime") {  // per-stage event times on the fused path (stage_report); lanes no longer overlap while it is on
        cudaDeviceSynchronize();
        for (Lane* L : lanes_) if (L->fast) L->fast->set_ktime(value);
        return 0;
    }
    if (key == "stage_pageable") { stage_pageable_ = value != 0; return 0; }
    if (key == "batch") { batch_ = value < 0 ? 0 : (value > V46_MAX_BATCH ? V46_MAX_BATCH : value); return 0; }
    if (key == "async") { async_ = value != 0; return 0; }
    if (key == "combine") { combine_ = value != 0; return 0; }
    if (key == "fuse") { cudaDeviceSynchronize(); for (Lane* L : lanes_) for (auto& r : L->run) if (r) { r->fuse = value != 0; r->clear_plans(); } return 0; }
    set_error("unknown option " + key);
    return -1;
}

std::string Engine::stage_report() {
    std::lock_guard<std::mutex> lk(mu_);
    std::string r;
    for (size_t i = 0; i < lanes_.size(); i++)
        if (lanes_[i]->fast) r += "lane\t" + std::to_string(i) + "\n" + lanes_[i]->fast->stage_report();
    return r;
}

void Engine::forget_frames() {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& f : frames_) f.host = nullptr;
}

int Engine::get_option(const std::string& key, int* value) {
    std::lock_guard<std::mutex> lk(mu_);
    if (key == "precision") *value = precision_;
    else if (key == "lanes") *value = (int)lanes_.size();
    else if (key == "fast") *value = use_fast_;
    else if (key == "batch") *value = batch_;
    else if (key == "plain_blocks") *value = plain_mask_;
    else if (key == "recompute_fm") *value = recompute_fm_;
    else if (key == "combine") *value = combine_;
    else if (key == "cpu_crop_quirk") *value = cpu_crop_quirk_;
    else if (key == "head_pack") *value = head_pack_;
    else if (key == "bgr") *value = bgr_;
    else if (key == "frame_cache") *value = frame_cache_;
    else if (key == "stage_pageable") *value = stage_pageable_;
    else if (key == "frame_cache_hits") *value = (int)frame_hits_;
    else if (key == "combined_batches") *value = (int)combiner_.batches();
    else if (key == "combined_requests") *value = (int)combiner_.requests();
    else if (key == "fast_active") *value = fast_usable();
    else { set_error("unknown option " + key); return -1; }
    return 0;
}

// The fused path is only trusted after it reproduced the generic executor (same weights, same device) on a small random
// frame pair: any model whose graph is not the expected rife-v4.6 IFNet fails init() or this check and keeps the
// generic path.
void Engine::setup_fast() {
    fast_ok_ = false;
    for (Lane* L : lanes_) { delete L->fast; L->fast = nullptr; }
    if (!v4_ || !loaded_ || !lanes_[0]->run[0]) return;
    std::string err;
    for (Lane* L : lanes_) {
        L->fast = new V46Runner();
        L->fast->set_recompute(recompute_fm_);
        L->fast->set_crop_quirk(cpu_crop_quirk_);
        L->fast->set_bgr(bgr_);
        L->fast->set_head_pack(head_pack_);
        if (L->fast->init(nets_[0].get(), lanes_[0]->run[0], err)) {
            for (Lane* L2 : lanes_) { delete L2->fast; L2->fast = nullptr; }
            return;
        }
    }
    const int w = 96, h = 64;
    const size_t n = (size_t)w * h * 3;
    std::vector<uint8_t> a(n), b(n), o0(n), o1(n);
    uint32_t s = 2463534242u;
    for (size_t i = 0; i < n; i++) {
        s = s * 1664525u + 1013904223u;
        int base = (int)(96 + 64 * sinf(0.11f * (float)((i / 3) % w)) + 40 * cosf(0.07f * (float)((i / 3) / w)));
        a[i] = (uint8_t)std::min(255, std::max(0, base + (int)((s >> 24) & 15)));
        b[i] = (uint8_t)std::min(255, std::max(0, base + 9 + (int)((s >> 16) & 15)));
    }
    Lane& L = *lanes_[0];
    DevBuf d0, d1, dout;
    if (d0.ensure(n) || d1.ensure(n) || dout.ensure(n)) return;
    cudaMemcpy(d0.p, a.data(), n, cudaMemcpyHostToDevice);
    cudaMemcpy(d1.p, b.data(), n, cudaMemcpyHostToDevice);
    const int saved_mode = L.run[0]->tc_mode;
    bool ok = false;
    do {
        // generic executor at the same precision tier (tensor cores, split operands)
        for (Lane* LL : lanes_) { LL->run[0]->tc_mode = 1; LL->run[0]->clear_plans(); }
        bool tta = tta_, ttat = ttat_;
        tta_ = ttat_ = false;
        int r = run_v4(L, d0.u8(), d1.u8(), w, h, 0.5f, dout.u8(), L.st);
        tta_ = tta; ttat_ = ttat;
        if (r || cudaStreamSynchronize(L.st) != cudaSuccess) break;
        cudaMemcpy(o0.data(), dout.p, n, cudaMemcpyDeviceToHost);
        if (L.fast->run(d0.u8(), d1.u8(), w, h, 0.5f, dout.u8(), L.st, err) || cudaStreamSynchronize(L.st) != cudaSuccess) break;
        cudaMemcpy(o1.data(), dout.p, n, cudaMemcpyDeviceToHost);
        int maxd = 0;
        size_t ne = 0;
        for (size_t i = 0; i < n; i++) { int d = abs((int)o0[i] - (int)o1[i]); maxd = std::max(maxd, d); ne += d != 0; }
        ok = maxd <= 1 && ne * 200 < n;  // identical up to fp32 rounding: at most a few 1-LSB flips
        if (getenv("RIFE_B200_VERBOSE")) fprintf(stderr, "[rife_b200] fused-path self-check (%s layout): max diff %d, %zu of %zu bytes differ -> %s\n", L.fast->is_v4() ? "rife-v4" : "rife-v4.6", maxd, ne, n, ok ? "enabled" : "disabled");
    } while (0);
    cudaGetLastError();
    for (Lane* LL : lanes_) { LL->run[0]->tc_mode = saved_mode; LL->run[0]->clear_plans(); }
    d0.release(); d1.release(); dout.release();
    fast_ok_ = ok;
    if (!ok) for (Lane* LL : lanes_) { delete LL->fast; LL->fast = nullptr; }
    else for (Lane* LL : lanes_) LL->fast->set_plain_mask(plain_mask_);
}

Tensor Engine::keep(const Tensor& t, DevBuf& b, cudaStream_t st) {
    Tensor o = t;
    b.ensure(t.count() * sizeof(float));
    cudaMemcpyAsync(b.p, t.p, t.count() * sizeof(float), cudaMemcpyDeviceToDevice, st);
    o.p = b.f();
    return o;
}

// pairs per lock-step batch: enough images to fill the machine in the coarse IFBlocks (about one 4K frame of pixels)
static int batch_rule(bool fast, int batch_opt, int w, int h) {
    if (!fast) return 1;
    if (batch_opt > 0) return batch_opt;
    const size_t px = (size_t)((w + 31) / 32 * 32) * ((h + 31) / 32 * 32);
    size_t b = (size_t)2 * 3840 * 2176 / (px ? px : 1);  // measured: 8 pairs at 1080p, 2 at 4K (profiles/README.md)
    if (b < 1) b = 1;
    if (b > V46_MAX_BATCH) b = V46_MAX_BATCH;
    return (int)b;
}
int Engine::batch_for(int w, int h) const { return batch_rule(fast_usable(), batch_, w, h); }

// n pairs on one lane: one lock-step batch on the fused path, otherwise pair by pair
int Engine::run_chunk(Lane& L, int n, const uint8_t* const* d_in0, const uint8_t* const* d_in1, int w, int h, const float* ts, uint8_t* const* d_out, cudaStream_t st) {
    if (n > 1 && fast_usable() && L.fast) {
        std::string err;
        int r = L.fast->run_batch(n, d_in0, d_in1, w, h, ts, d_out, st, err);
        if (r) { set_error(err); return -5; }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error(std::string("kernel launch failure: ") + cudaGetErrorString(e)); return -2; }
        return 0;
    }
    for (int i = 0; i < n; i++) {
        int r = run_device(L, d_in0[i], d_in1[i], w, h, ts[i], d_out[i], st);
        if (r) return r;
    }
    return 0;
}

uint8_t* Engine::stage_acquire(size_t bytes, int* idx) {
    std::lock_guard<std::mutex> lk(stage_mu_);
    int free_i = -1;
    for (size_t i = 0; i < stage_.size(); i++)
        if (!stage_[i].busy) { if (stage_[i].cap >= bytes) { stage_[i].busy = true; *idx = (int)i; return stage_[i].p; } if (free_i < 0) free_i = (int)i; }
    if (free_i < 0) {
        if (stage_.size() >= 64) return nullptr;  // more concurrent callers than that: let the driver stage
        stage_.emplace_back();
        free_i = (int)stage_.size() - 1;
    }
    StageSlot& s = stage_[free_i];
    if (s.p) cudaFreeHost(s.p);
    s.p = nullptr;
    s.cap = 0;
    cudaSetDevice(gpuid_);
    if (cudaHostAlloc((void**)&s.p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); s.p = nullptr; return nullptr; }
    s.cap = bytes;
    s.busy = true;
    *idx = free_i;
    return s.p;
}
void Engine::stage_release(int idx) {
    std::lock_guard<std::mutex> lk(stage_mu_);
    stage_[idx].busy = false;
}

static bool is_pageable(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

int Engine::process_host(const uint8_t* in0, const uint8_t* in1, int w, int h, float t, uint8_t* out) {
    const uint8_t* a[1] = {in0};
    const uint8_t* b[1] = {in1};
    uint8_t* o[1] = {out};
    if (!in0 || !in1 || !out) { set_error("bad argument"); return -1; }
    if (w > 0 && h > 0 && t != 0.f && t != 1.f && snap_stage_.load(std::memory_order_relaxed) && (is_pageable(in0) || is_pageable(in1) || is_pageable(out))) {
        // pageable caller memory: stage through a pinned slot in THIS thread (see engine.h), then run the same path on the slot
        const size_t nb = (size_t)w * h * 3;
        int idx = -1;
        if (uint8_t* sl = stage_acquire(3 * nb, &idx)) {
            memcpy(sl, in0, nb);
            memcpy(sl + nb, in1, nb);
            const int r = process_host(sl, sl + nb, w, h, t, sl + 2 * nb);
            if (!r) memcpy(out, sl + 2 * nb, nb);
            stage_release(idx);
            return r;
        }
    }
    // The reference's CLI calls process() from several proc threads on one object (src/main.cpp:346-366): requests that
    // arrive while another is being served are executed together as one lock-step batch instead of one after the other.
    // Decided from a lock-free snapshot: the leader of a running batch holds mu_, followers must still be able to queue.
    if (snap_combine_.load(std::memory_order_relaxed) && w > 0 && h > 0) {
        const int nl = snap_lanes_.load(std::memory_order_acquire);
        const int B = batch_rule(snap_fast_.load(std::memory_order_relaxed) != 0, snap_batch_.load(std::memory_order_relaxed), w, h);
        if (B > 1) {
            HostReq r = {in0, in1, w, h, t, out};
            return combiner_.submit(&r, B * nl, [this](HostReq** rq, int n) { return run_combined(rq, n); });
        }
    }
    return process_batch(1, a, b, w, h, &t, o);
}

// n queued process() calls: one process_batch per run of equal frame size (a real caller has a single size)
int Engine::run_combined(HostReq** rq, int n) {
    const uint8_t* a[Combiner<HostReq>::kMax];
    const uint8_t* b[Combiner<HostReq>::kMax];
    uint8_t* o[Combiner<HostReq>::kMax];
    float ts[Combiner<HostReq>::kMax];
    int rc = 0;
    for (int i = 0; i < n;) {
        int j = i;
        for (; j < n && rq[j]->w == rq[i]->w && rq[j]->h == rq[i]->h; j++) {
            a[j - i] = rq[j]->in0; b[j - i] = rq[j]->in1; o[j - i] = rq[j]->out; ts[j - i] = rq[j]->t;
        }
        const int r = process_batch(j - i, a, b, rq[i]->w, rq[i]->h, ts, o);
        if (r) rc = r;
        i = j;
    }
    return rc;
}

int Engine::process_device(const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out) {
    const uint8_t* a[1] = {d_in0};
    const uint8_t* b[1] = {d_in1};
    uint8_t* o[1] = {d_out};
    if (!d_in0 || !d_in1 || !d_out) { set_error("bad argument"); return -1; }
    return process_batch_device(1, a, b, w, h, &t, o);
}

// Frames already in device memory.  Pairs are cut into chunks (lock-step batches on the fused path) that are dealt
// round-robin to the lanes; when the caller supplied a stream (set_stream) every lane first waits for that stream and
// the stream finally waits for every lane, so events the caller records on it bracket all of the work.
int Engine::process_batch_device(int n, const uint8_t* const* d_in0, const uint8_t* const* d_in1, int w, int h, const float* ts, uint8_t* const* d_out) {
    if (n < 0 || !d_in0 || !d_in1 || !d_out || !ts || w <= 0 || h <= 0) { set_error("bad argument"); return -1; }
    for (int i = 0; i < n; i++)  // validate everything before anything is queued
        if (!d_in0[i] || !d_in1[i] || !d_out[i]) { set_error("null frame pointer"); return -1; }
    size_t nb = (size_t)w * h * 3;
    std::lock_guard<std::mutex> lk(mu_);
    if (!loaded_) { set_error("process before load"); return -4; }
    cudaSetDevice(gpuid_);
    const int nl = tta_ ? 1 : (int)lanes_.size();  // spatial TTA: lane 0 coordinates, the other lanes are its helpers (tta_fork / tta_join)
    const int B = batch_for(w, h);
    if (use_user_stream_) {
        cudaEventRecord(ev_entry_, user_stream_);
        for (int l = 0; l < nl; l++) cudaStreamWaitEvent(lanes_[l]->st, ev_entry_, 0);
    }
    const uint8_t* c0[V46_MAX_BATCH];
    const uint8_t* c1[V46_MAX_BATCH];
    uint8_t* co[V46_MAX_BATCH];
    float ct[V46_MAX_BATCH];
    int cn = 0, chunk = 0;
    // spread the work: no more pairs per chunk than needed to give every lane something to do
    int per = B;
    if (n < per * nl) per = (n + nl - 1) / nl;
    if (per < 1) per = 1;
    for (int i = 0; i <= n; i++) {
        if (i < n) {
            if (ts[i] == 0.f || ts[i] == 1.f) {  // rife.cpp:3206-3216: the output is an input
                cudaMemcpyAsync(d_out[i], ts[i] == 0.f ? d_in0[i] : d_in1[i], nb, cudaMemcpyDeviceToDevice, lanes_[chunk % nl]->st);
                continue;
            }
            c0[cn] = d_in0[i]; c1[cn] = d_in1[i]; co[cn] = d_out[i]; ct[cn] = ts[i];
            cn++;
        }
        if (cn == per || (i == n && cn > 0)) {
            Lane& L = *lanes_[chunk % nl];
            int r = run_chunk(L, cn, c0, c1, w, h, ct, co, L.st);
            if (r) { sync_all(); return r; }  // nothing of this call may still be running when the caller sees the error
            cn = 0;
            chunk++;
        }
    }
    for (int l = 0; l < nl; l++) {
        cudaEventRecord(lanes_[l]->done, lanes_[l]->st);
        if (use_user_stream_) cudaStreamWaitEvent(user_stream_, lanes_[l]->done, 0);
    }
    if (async_) return 0;
    cudaError_t e = cudaSuccess;
    for (int l = 0; l < nl; l++) { cudaError_t e2 = cudaStreamSynchronize(lanes_[l]->st); if (e2 != cudaSuccess) e = e2; }
    if (use_user_stream_) { cudaError_t e2 = cudaStreamSynchronize(user_stream_); if (e2 != cudaSuccess) e = e2; }
    if (e != cudaSuccess) { set_error(std::string("CUDA failure: ") + cudaGetErrorString(e)); return -2; }
    return 0;
}

// Device copy of the host frame `host`: an entry that already holds it (same pointer and size; uploaded earlier in this
// call, or in an earlier call when option "frame_cache" is on), else the least recently used entry that the chunk being
// assembled (`cur`, `ncur`) does not use.  An evicted entry may still be read by a chunk in flight: the caller orders the new
// upload behind its read_done event.  *hit tells the caller whether an upload is needed.
Engine::FrameEntry* Engine::frame_lookup(const uint8_t* host, size_t nb, FrameEntry* const* cur, int ncur, bool* hit) {
    const size_t kMaxFrames = 4 * (size_t)V46_MAX_BATCH * 2 + 2;  // ~4 chunks of 8 pairs with all-distinct frames
    if (frames_.capacity() < kMaxFrames) frames_.reserve(kMaxFrames);  // entries are referenced by pointer: never reallocate
    // order of preference on a miss: an entry nobody owns (a call that re-sends few frames keeps using the same few buffers), a
    // new entry, the least recently used one outside the current chunk
    FrameEntry* lru = nullptr;
    FrameEntry* spare = nullptr;
    for (auto& f : frames_) {
        if (f.host == host && f.nb == nb) { *hit = true; f.stamp = ++frame_clock_; return &f; }
        if (!f.host) { if (!spare) spare = &f; continue; }
        bool in_cur = false;
        for (int u = 0; u < ncur; u++) in_cur = in_cur || cur[u] == &f;
        if (!in_cur && (!lru || f.stamp < lru->stamp)) lru = &f;
    }
    *hit = false;
    if (spare) lru = spare;
    else if (frames_.size() < kMaxFrames) {
        frames_.emplace_back();
        lru = &frames_.back();
        if (cudaEventCreateWithFlags(&lru->read_done, cudaEventDisableTiming) != cudaSuccess) { frames_.pop_back(); return nullptr; }
    }
    if (!lru) return nullptr;
    lru->host = host;
    lru->nb = nb;
    lru->stamp = ++frame_clock_;
    return lru;
}

// A larger frame size than the table's buffers were made for: drop the small buffers in one go.  (Growing them one by one as
// the LRU order reaches them puts a device-wide cudaFree into the pipeline of the next several calls.)
void Engine::frames_fit(size_t nb) {
    bool small = false;
    for (auto& f : frames_) small = small || (f.buf.cap && f.buf.cap < nb);
    if (!small) return;
    sync_all();
    for (auto& f : frames_)
        if (f.buf.cap && f.buf.cap < nb) { f.buf.release(); f.host = nullptr; f.reading = false; }
}

// Host frames, pipelined by chunk: H2D runs on one copy stream, compute on lane chunk mod lanes, D2H on the other copy
// stream, chained with events (pinned host memory makes the copies truly asynchronous; pageable memory still works, the
// driver then stages synchronously).  Input frames live in the frame table (frame_lookup); outputs in per-slot buffers.
int Engine::process_batch(int n, const uint8_t* const* in0, const uint8_t* const* in1, int w, int h, const float* ts, uint8_t* const* out) {
    if (n < 0 || !in0 || !in1 || !out || !ts || w <= 0 || h <= 0) { set_error("bad argument"); return -1; }
    for (int i = 0; i < n; i++)  // validate everything before anything is queued
        if (!in0[i] || !in1[i] || !out[i]) { set_error("null frame pointer"); return -1; }
    size_t nb = (size_t)w * h * 3;
    std::lock_guard<std::mutex> lk(mu_);
    if (!loaded_) { set_error("process before load"); return -4; }
    cudaSetDevice(gpuid_);
    const int nl = tta_ ? 1 : (int)lanes_.size();  // spatial TTA: lane 0 coordinates, the other lanes are its helpers
    const int nslots = 2 * nl <= kSlots ? 2 * nl : kSlots;
    const int B = batch_for(w, h);
    int per = B;
    if (n < per * nl) per = (n + nl - 1) / nl;
    if (per < 1) per = 1;
    if (out_u8_.size() < (size_t)kSlots * V46_MAX_BATCH) out_u8_.resize((size_t)kSlots * V46_MAX_BATCH);
    frames_fit(nb);
    std::vector<int> used(nslots, 0);
    const uint8_t* c0[V46_MAX_BATCH];
    const uint8_t* c1[V46_MAX_BATCH];
    uint8_t* co[V46_MAX_BATCH];
    uint8_t* ho[V46_MAX_BATCH];
    float ct[V46_MAX_BATCH];
    FrameEntry* fe[2 * V46_MAX_BATCH];
    int nfe = 0;
    int cn = 0, chunk = 0;
    int rc = 0;
    // RIFE_B200_TRACE_BATCH=1: per chunk, CUDA-event times of compute start / end and of the end of its D2H copies (stderr)
    static const bool trace = getenv("RIFE_B200_TRACE_BATCH") && atoi(getenv("RIFE_B200_TRACE_BATCH")) != 0;
    struct Tr { cudaEvent_t a, b, c; int lane; };
    std::vector<Tr> tr;
    const auto host_t0 = std::chrono::steady_clock::now();
    std::vector<double> host_issue;
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    double tr_first[4] = {0, 0, 0, 0};  // chunk 0: host time after the uploads are queued / the waits / the kernels / the downloads
    for (int i = 0; i <= n && !rc; i++) {
        if (i < n) {
            if (ts[i] == 0.f || ts[i] == 1.f) {
                if (out[i] != (ts[i] == 0.f ? in0[i] : in1[i])) memcpy(out[i], ts[i] == 0.f ? in0[i] : in1[i], nb);  // host-side copy, touches nothing queued
                continue;
            }
            const int s = chunk % nslots;
            if (cn == 0) nfe = 0;
            const uint8_t* hp2[2] = {in0[i], in1[i]};
            const uint8_t* dp2[2] = {nullptr, nullptr};
            for (int k = 0; k < 2 && !rc; k++) {
                bool hit = false;
                FrameEntry* f = frame_lookup(hp2[k], nb, fe, nfe, &hit);
                if (!f || (!hit && f->buf.ensure(nb))) { set_error("cudaMalloc failed"); rc = -2; break; }
                if (!hit) {
                    if (f->reading) cudaStreamWaitEvent(st_copy_[0], f->read_done, 0);  // the previous content is still being read by a lane
                    f->reading = false;
                    cudaMemcpyAsync(f->buf.p, hp2[k], nb, cudaMemcpyHostToDevice, st_copy_[0]);
                    g_h2d_bytes += nb;
                } else frame_hits_++;
                bool listed = false;
                for (int u = 0; u < nfe; u++) listed = listed || fe[u] == f;
                if (!listed) fe[nfe++] = f;
                dp2[k] = f->buf.u8();
            }
            if (rc) break;
            DevBuf& b2 = out_u8_[(size_t)s * V46_MAX_BATCH + cn];
            if (b2.ensure(nb)) { set_error("cudaMalloc failed"); rc = -2; break; }
            c0[cn] = dp2[0]; c1[cn] = dp2[1]; co[cn] = b2.u8(); ho[cn] = out[i]; ct[cn] = ts[i];
            cn++;
        }
        if (cn == per || (i == n && cn > 0)) {
            const int s = chunk % nslots;
            Lane& L = *lanes_[chunk % nl];
            if (trace && chunk == 0) tr_first[0] = host_ms();
            cudaEventRecord(ev_h2d_[s], st_copy_[0]);
            cudaStreamWaitEvent(L.st, ev_h2d_[s], 0);
            if (used[s]) cudaStreamWaitEvent(L.st, ev_d2h_[s], 0);  // output buffers of slot s have been downloaded
            if (trace) { Tr t; cudaEventCreate(&t.a); cudaEventCreate(&t.b); cudaEventCreate(&t.c); t.lane = chunk % nl; cudaEventRecord(t.a, L.st); tr.push_back(t); }
            if (trace && chunk == 0) tr_first[1] = host_ms();
            int r = run_chunk(L, cn, c0, c1, w, h, ct, co, L.st);
            if (r) { rc = r; break; }
            if (trace && chunk == 0) tr_first[2] = host_ms();
            if (trace) cudaEventRecord(tr.back().b, L.st);
            cudaEventRecord(ev_comp_[s], L.st);
            for (int u = 0; u < nfe; u++) { cudaEventRecord(fe[u]->read_done, L.st); fe[u]->reading = true; }
            // results go home on two dedicated copy streams (default), or on the lane's own stream (RIFE_B200_D2H=1: no cross-stream
            // hand-over, but the lane's next chunk then waits for the copy; never with a single lane)
            const bool on_lane = d2h_on_lane_ && nl > 1;
            cudaStream_t sd = on_lane ? L.st : st_copy_[1 + (chunk & 1)];
            if (!on_lane) cudaStreamWaitEvent(sd, ev_comp_[s], 0);
            for (int k = 0; k < cn; k++) cudaMemcpyAsync(ho[k], co[k], nb, cudaMemcpyDeviceToHost, sd);
            g_d2h_bytes += (unsigned long long)cn * nb;
            cudaEventRecord(ev_d2h_[s], sd);
            if (trace) { cudaEventRecord(tr.back().c, sd); host_issue.push_back(host_ms()); }
            used[s] = 1;
            cn = 0;
            chunk++;
        }
    }
    // everything queued by this call finishes before it returns -- also on the error paths, where the caller is about to
    // reuse or free its buffers
    cudaError_t e = cudaStreamSynchronize(st_copy_[0]);
    for (Lane* L : lanes_) { cudaError_t e2 = cudaStreamSynchronize(L->st); if (e2 != cudaSuccess) e = e2; }
    for (int k = 1; k < 3; k++) { cudaError_t e3 = cudaStreamSynchronize(st_copy_[k]); if (e3 != cudaSuccess) e = e3; }
    if (trace && !tr.empty()) {
        fprintf(stderr, "[rife_b200 trace] process_batch n=%d: chunk lane | compute start..end | d2h end (ms since the first chunk's start) | host issued at\n", n);
        for (size_t i = 0; i < tr.size(); i++) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, tr[0].a, tr[i].a); cudaEventElapsedTime(&b, tr[0].a, tr[i].b); cudaEventElapsedTime(&c, tr[0].a, tr[i].c);
            fprintf(stderr, "[rife_b200 trace] %2zu %d | %7.2f .. %7.2f | %7.2f | %7.2f\n", i, tr[i].lane, a, b, c, i < host_issue.size() ? host_issue[i] : -1.0);
        }
        for (auto& t : tr) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); cudaEventDestroy(t.c); }
        cudaGetLastError();
        fprintf(stderr, "[rife_b200 trace] chunk 0 on the host: uploads queued %.2f, waits queued %.2f, kernels queued %.2f ms\n", tr_first[0], tr_first[1], tr_first[2]);
        fprintf(stderr, "[rife_b200 trace] call returned after %.2f ms on the host\n", host_ms());
    }
    for (auto& f : frames_) { f.reading = false; if (!frame_cache_ || rc) f.host = nullptr; }
    if (rc) { cudaGetLastError(); return rc; }
    if (e != cudaSuccess) { set_error(std::string("CUDA failure: ") + cudaGetErrorString(e)); return -2; }
    return 0;
}

int Engine::run_device(Lane& L, const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out, cudaStream_t st) {
    if (fast_usable() && L.fast) {
        std::string err;
        int r = L.fast->run(d_in0, d_in1, w, h, t, d_out, st, err);
        if (r) { set_error(err); return -5; }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error(std::string("kernel launch failure: ") + cudaGetErrorString(e)); return -2; }
        return 0;
    }
    int r = v4_ ? run_v4(L, d_in0, d_in1, w, h, t, d_out, st) : run_v1v2(L, d_in0, d_in1, w, h, d_out, st);
    if (r) return r;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("kernel launch failure: ") + cudaGetErrorString(e)); return -2; }
    return 0;
}

typedef std::vector<std::pair<std::string, Tensor>> Inputs;

// TTA modes: the 8 orientations (x 2 time directions) of one pair are independent graph walks between the flow-averaging
// points (rife.cpp:1541-1949, 3432-3823).  The lane that serves the pair (the coordinator) deals them round-robin to ALL lanes
// of the engine -- each with its own stream, plans and arena -- so the many small kernels of different orientations overlap;
// fork() / join() order the helpers' streams after / before the coordinator's.  (process_batch gives TTA pairs to lane 0 only.)
void Engine::tta_fork(Lane& L) {
    if (!tta_) return;
    cudaEventRecord(L.fork, L.st);
    for (Lane* H : lanes_) if (H != &L) cudaStreamWaitEvent(H->st, L.fork, 0);
}
void Engine::tta_join(Lane& L) {
    if (!tta_) return;
    for (Lane* H : lanes_) if (H != &L) { cudaEventRecord(H->done, H->st); cudaStreamWaitEvent(L.st, H->done, 0); }
}
Lane& Engine::tta_lane(Lane& L, int job) { return tta_ ? *lanes_[(size_t)job % lanes_.size()] : L; }

// ---- rife-v4 / v4.6: rife.cpp:3204-4401 -------------------------------------------------------------------
int Engine::run_v4(Lane& L, const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, float t, uint8_t* d_out, cudaStream_t st) {
    const int wp = (w + 31) / 32 * 32, hp = (h + 31) / 32 * 32;  // rife.cpp:3229-3230
    const size_t plane = (size_t)wp * hp;
    NetRunner& F = *L.run[0];
    std::string err;
    std::vector<Tensor> o;
    const int nti = tta_ ? 8 : 1;
    Tensor I0[8], I1[8], T[2], TR[2];
    if (L.pad0.ensure((size_t)nti * 3 * plane * 4) || L.pad1.ensure((size_t)nti * 3 * plane * 4)) { set_error("cudaMalloc failed"); return -2; }
    launch_preproc(d_in0, w, h, L.pad0.f(), wp, hp, nti, bgr_, st);  // rife.cpp:4152-4211 / 3253-3413: all orientations from one read
    launch_preproc(d_in1, w, h, L.pad1.f(), wp, hp, nti, bgr_, st);
    for (int ti = 0; ti < nti; ti++) {
        int th = ti < 4 ? hp : wp, tw = ti < 4 ? wp : hp;
        I0[ti] = Tensor::chw(L.pad0.f() + (size_t)ti * 3 * plane, 3, th, tw);
        I1[ti] = Tensor::chw(L.pad1.f() + (size_t)ti * 3 * plane, 3, th, tw);
    }
    L.ts[0].ensure(plane * 4);
    launch_fill(L.ts[0].f(), plane, t, st);  // full padded plane, rife.cpp:4213-4214
    T[0] = Tensor::chw(L.ts[0].f(), 1, hp, wp);
    T[1] = Tensor::chw(L.ts[0].f(), 1, wp, hp);  // rife.cpp:3313-3316 (same constant, transposed extent)
    if (ttat_) {
        L.tsr[0].ensure(plane * 4);
        launch_fill(L.tsr[0].f(), plane, 1.f - t, st);
        TR[0] = Tensor::chw(L.tsr[0].f(), 1, hp, wp);
        TR[1] = Tensor::chw(L.tsr[0].f(), 1, wp, hp);
    }
    static const char* kFlow[4] = {"flow0", "flow1", "flow2", "flow3"};

    if (!tta_ && !ttat_) {
        // rife.cpp:4345-4351
        Inputs in = {{"in0", I0[0]}, {"in1", I1[0]}, {"in2", T[0]}};
        if (F.run(in, {"out0"}, o, st, err)) { set_error(err); return -5; }
        const float* ins[1] = {o[0].p};
        launch_postproc(ins, 1, wp, hp, d_out, w, h, cpu_crop_quirk_, bgr_, st);  // rife.cpp:4375-4398
        return 0;
    }

    Tensor fl[4][8], flr[4][8];
    for (int fi = 0; fi < 4; fi++) {
        tta_fork(L);
        for (int ti = 0; ti < nti; ti++) {
            Lane& H = tta_lane(L, ti);
            NetRunner& FH = *H.run[0];
            {   // rife.cpp:3432-3451 / 4233-4252: inject the already merged flow0..fi-1, extract flow<fi>
                Inputs in = {{"in0", I0[ti]}, {"in1", I1[ti]}, {"in2", T[ti / 4]}};
                for (int k = 0; k < fi; k++) in.push_back({kFlow[k], fl[k][ti]});
                if (FH.run(in, {kFlow[fi]}, o, H.st, err)) { set_error(err); return -5; }
                fl[fi][ti] = keep(o[0], L.flow[fi][ti], H.st);
            }
            if (ttat_) {
                Inputs in = {{"in0", I1[ti]}, {"in1", I0[ti]}, {"in2", TR[ti / 4]}};
                for (int k = 0; k < fi; k++) in.push_back({kFlow[k], flr[k][ti]});
                if (FH.run(in, {kFlow[fi]}, o, H.st, err)) { set_error(err); return -5; }
                flr[fi][ti] = keep(o[0], L.flowr[fi][ti], H.st);
                // rife.cpp:3476-3512 / 4277-4312
                launch_temporal_merge_v2(fl[fi][ti].p, flr[fi][ti].p, (size_t)fl[fi][ti].h * fl[fi][ti].w, 1, H.st);
            }
        }
        tta_join(L);
        if (tta_) {  // rife.cpp:3515-3668 (+ reversed set :3670-3823)
            float* f8[8];
            for (int ti = 0; ti < 8; ti++) f8[ti] = fl[fi][ti].p;
            launch_flow_tta_avg(f8, 5, fl[fi][0].w, fl[fi][0].h, st);
            if (ttat_) {
                for (int ti = 0; ti < 8; ti++) f8[ti] = flr[fi][ti].p;
                launch_flow_tta_avg(f8, 5, fl[fi][0].w, fl[fi][0].h, st);
            }
        }
    }
    const float* ins[16];
    tta_fork(L);
    for (int ti = 0; ti < nti; ti++) {
        Lane& H = tta_lane(L, ti);
        NetRunner& FH = *H.run[0];
        Inputs in = {{"in0", I0[ti]}, {"in1", I1[ti]}, {"in2", T[ti / 4]}};
        for (int k = 0; k < 4; k++) in.push_back({kFlow[k], fl[k][ti]});
        if (FH.run(in, {"out0"}, o, H.st, err)) { set_error(err); return -5; }
        ins[ti] = keep(o[0], L.outp[ti], H.st).p;
        if (ttat_) {
            Inputs inr = {{"in0", I1[ti]}, {"in1", I0[ti]}, {"in2", TR[ti / 4]}};
            for (int k = 0; k < 4; k++) inr.push_back({kFlow[k], flr[k][ti]});
            if (FH.run(inr, {"out0"}, o, H.st, err)) { set_error(err); return -5; }
            ins[nti + ti] = keep(o[0], L.outp[8 + ti], H.st).p;
        }
    }
    tta_join(L);
    if (tta_) launch_postproc(ins, ttat_ ? 16 : 8, wp, hp, d_out, w, h, 0, bgr_, st);  // rife.cpp:4060-4144
    else launch_postproc(ins, 2, wp, hp, d_out, w, h, cpu_crop_quirk_, bgr_, st);       // rife.cpp:4356-4371
    return 0;
}

// ---- rife / rife-HD / rife-UHD / rife-anime (v1), rife-v2.x / v3.x (v2): rife.cpp:1214-2460 ----------------
int Engine::run_v1v2(Lane& L, const uint8_t* d_in0, const uint8_t* d_in1, int w, int h, uint8_t* d_out, cudaStream_t st) {
    const int wp = (w + 31) / 32 * 32, hp = (h + 31) / 32 * 32;
    const size_t plane = (size_t)wp * hp;
    std::string err;
    std::vector<Tensor> o;
    const int nti = tta_ ? 8 : 1;
    Tensor I0[8], I1[8];
    if (L.pad0.ensure((size_t)nti * 3 * plane * 4) || L.pad1.ensure((size_t)nti * 3 * plane * 4)) { set_error("cudaMalloc failed"); return -2; }
    launch_preproc(d_in0, w, h, L.pad0.f(), wp, hp, nti, bgr_, st);  // rife.cpp:4152-4211 / 3253-3413: all orientations from one read
    launch_preproc(d_in1, w, h, L.pad1.f(), wp, hp, nti, bgr_, st);
    for (int ti = 0; ti < nti; ti++) {
        int th = ti < 4 ? hp : wp, tw = ti < 4 ? wp : hp;
        I0[ti] = Tensor::chw(L.pad0.f() + (size_t)ti * 3 * plane, 3, th, tw);
        I1[ti] = Tensor::chw(L.pad1.f() + (size_t)ti * 3 * plane, 3, th, tw);
    }
    // flownet(a, b) -> flow at half resolution, on lane H (its executor, stream and scratch); uhd: rife.cpp:2212-2229
    auto flownet = [&](Lane& H, const Tensor& a, const Tensor& b, DevBuf& dst, Tensor& flow) -> int {
        NetRunner& FH = *H.run[0];
        cudaStream_t sh = H.st;
        if (uhd_) {
            Tensor ad = Tensor::chw(nullptr, 3, (int)(a.h * 0.5f), (int)(a.w * 0.5f)), bd = ad;
            H.tmp[0].ensure(ad.count() * 4);
            H.tmp[1].ensure(ad.count() * 4);
            ad.p = H.tmp[0].f();
            bd.p = H.tmp[1].f();
            launch_interp_bilinear(a.p, 3, a.h, a.w, ad.p, ad.h, ad.w, sh);
            launch_interp_bilinear(b.p, 3, b.h, b.w, bd.p, bd.h, bd.w, sh);
            Inputs in = {{"input0", ad}, {"input1", bd}};
            if (FH.run(in, {"flow"}, o, sh, err)) { set_error(err); return -5; }
            Tensor fd = o[0];
            flow = Tensor::chw(nullptr, fd.c, (int)(fd.h * 2.f), (int)(fd.w * 2.f));
            dst.ensure(flow.count() * 4);
            flow.p = dst.f();
            launch_interp_bilinear(fd.p, fd.c, fd.h, fd.w, flow.p, flow.h, flow.w, sh);
            launch_unary(flow.p, flow.p, flow.count(), U_MUL_S, 2.f, 0.f, sh);
        } else {
            Inputs in = {{"input0", a}, {"input1", b}};
            if (FH.run(in, {"flow"}, o, sh, err)) { set_error(err); return -5; }
            flow = keep(o[0], dst, sh);
        }
        return 0;
    };
    Tensor fl[8], flr[8];
    auto merge = [&](int ti, cudaStream_t sm) {
        size_t n = (size_t)fl[ti].h * fl[ti].w;
        if (v2_) launch_temporal_merge_v2(fl[ti].p, flr[ti].p, n, 0, sm);  // rife.cpp:2285-2306
        else launch_temporal_merge_v1(fl[ti].p, flr[ti].p, n, sm);         // rife.cpp:2307-2319
    };
    tta_fork(L);
    for (int ti = 0; ti < nti; ti++) {
        Lane& H = tta_lane(L, ti);
        if (flownet(H, I0[ti], I1[ti], L.flow[0][ti], fl[ti])) return -5;
        if (ttat_) {
            if (flownet(H, I1[ti], I0[ti], L.flowr[0][ti], flr[ti])) return -5;
            merge(ti, H.st);
        }
    }
    tta_join(L);
    if (tta_) {  // rife.cpp:1541-1719, reversed :1721-1896, second merge :1898-1949
        float* f8[8];
        for (int ti = 0; ti < 8; ti++) f8[ti] = fl[ti].p;
        launch_flow_tta_avg(f8, v2_ ? 4 : 2, fl[0].w, fl[0].h, st);
        if (ttat_) {
            for (int ti = 0; ti < 8; ti++) f8[ti] = flr[ti].p;
            launch_flow_tta_avg(f8, v2_ ? 4 : 2, fl[0].w, fl[0].h, st);
            for (int ti = 0; ti < 8; ti++) merge(ti, st);
        }
    }
    static const char* kCtx[4] = {"f1", "f2", "f3", "f4"};
    const float* ins[16];
    tta_fork(L);
    for (int ti = 0; ti < nti; ti++) {
        Lane& H = tta_lane(L, ti);
        NetRunner& CH = *H.run[1];
        NetRunner& UH = *H.run[2];
        cudaStream_t sh = H.st;
        Tensor c0[4], c1[4];
        Tensor f0 = fl[ti], f1 = fl[ti];
        if (v2_) {  // Slice 4 -> 2 + 2, rife.cpp:2322-2330
            f0.c = 2;
            f1.c = 2;
            f1.p = fl[ti].p + 2 * (size_t)fl[ti].h * fl[ti].w;
        }
        {   // rife.cpp:2335-2351
            Inputs in = {{"input.1", I0[ti]}, {"flow.0", f0}};
            if (CH.run(in, {kCtx[0], kCtx[1], kCtx[2], kCtx[3]}, o, sh, err)) { set_error(err); return -5; }
            for (int k = 0; k < 4; k++) c0[k] = keep(o[k], H.ctx[0][k], sh);
        }
        {   // rife.cpp:2352-2368
            Inputs in = {{"input.1", I1[ti]}, {v2_ ? "flow.0" : "flow.1", f1}};
            if (CH.run(in, {kCtx[0], kCtx[1], kCtx[2], kCtx[3]}, o, sh, err)) { set_error(err); return -5; }
            for (int k = 0; k < 4; k++) c1[k] = keep(o[k], H.ctx[1][k], sh);
        }
        {   // rife.cpp:2372-2388
            Inputs in = {{"img0", I0[ti]}, {"img1", I1[ti]}, {"flow", fl[ti]}, {"3", c0[0]}, {"4", c0[1]}, {"5", c0[2]}, {"6", c0[3]},
                         {"7", c1[0]}, {"8", c1[1]}, {"9", c1[2]}, {"10", c1[3]}};
            if (UH.run(in, {"output"}, o, sh, err)) { set_error(err); return -5; }
            ins[ti] = (tta_ || ttat_) ? keep(o[0], L.outp[ti], sh).p : o[0].p;
        }
        if (ttat_) {  // rife.cpp:2391-2409
            Inputs in = {{"img0", I1[ti]}, {"img1", I0[ti]}, {"flow", flr[ti]}, {"3", c1[0]}, {"4", c1[1]}, {"5", c1[2]}, {"6", c1[3]},
                         {"7", c0[0]}, {"8", c0[1]}, {"9", c0[2]}, {"10", c0[3]}};
            if (UH.run(in, {"output"}, o, sh, err)) { set_error(err); return -5; }
            ins[nti + ti] = keep(o[0], L.outp[8 + ti], sh).p;
        }
    }
    tta_join(L);
    if (tta_) launch_postproc(ins, ttat_ ? 16 : 8, wp, hp, d_out, w, h, 0, bgr_, st);
    else launch_postproc(ins, ttat_ ? 2 : 1, wp, hp, d_out, w, h, cpu_crop_quirk_, bgr_, st);
    return 0;
}

}  // namespace rife
