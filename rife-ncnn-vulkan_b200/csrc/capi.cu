// capi.cu -- the C ABI declared in include/rife_b200.h (thin shell over rife::Engine; catches everything).
#include <new>
#include <string>

#include "../../include/rife_b200.h"
#include "engine.h"
#include "kernels.h"

struct rife_b200 {
    rife::Engine* eng;
    std::string err;
};

#define GUARD_BEGIN try {
#define GUARD_END                                   \
    }                                               \
    catch (const std::bad_alloc&) {                 \
        return RIFE_B200_ERR_INTERNAL;              \
    }                                               \
    catch (...) {                                   \
        return RIFE_B200_ERR_INTERNAL;              \
    }

static int map_err(int r) {
    switch (r) {
        case 0: return RIFE_B200_OK;
        case -1: return RIFE_B200_ERR_ARG;
        case -2: return RIFE_B200_ERR_DEVICE;
        case -3: return RIFE_B200_ERR_MODEL;
        case -4: return RIFE_B200_ERR_STATE;
        default: return RIFE_B200_ERR_INTERNAL;
    }
}

extern "C" {

int rife_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rife_b200_create(rife_b200_t** handle, int gpuid, int tta_mode, int tta_temporal_mode, int uhd_mode, int /*num_threads*/, int rife_v2, int rife_v4) {
    GUARD_BEGIN
    if (!handle) return RIFE_B200_ERR_ARG;
    *handle = nullptr;
    if (gpuid < 0) return RIFE_B200_ERR_ARG;  // -1 = the reference's CPU mode: rejected by design
    rife_b200* h = new rife_b200();
    h->eng = new rife::Engine(gpuid, tta_mode != 0, tta_temporal_mode != 0, uhd_mode != 0, rife_v2 != 0, rife_v4 != 0);
    int r = h->eng->init();
    if (r) {
        delete h->eng;
        delete h;
        return map_err(r);
    }
    *handle = h;
    return RIFE_B200_OK;
    GUARD_END
}

int rife_b200_load(rife_b200_t* h, const char* modeldir) {
    GUARD_BEGIN
    if (!h || !modeldir) return RIFE_B200_ERR_ARG;
    return map_err(h->eng->load(modeldir));
    GUARD_END
}

int rife_b200_process(rife_b200_t* h, const unsigned char* in0, const unsigned char* in1, int w, int hh, float t, unsigned char* out) {
    GUARD_BEGIN
    if (!h) return RIFE_B200_ERR_ARG;
    return map_err(h->eng->process_host(in0, in1, w, hh, t, out));
    GUARD_END
}

int rife_b200_process_device(rife_b200_t* h, const unsigned char* in0, const unsigned char* in1, int w, int hh, float t, unsigned char* out) {
    GUARD_BEGIN
    if (!h) return RIFE_B200_ERR_ARG;
    return map_err(h->eng->process_device(in0, in1, w, hh, t, out));
    GUARD_END
}

int rife_b200_process_batch(rife_b200_t* h, int n, const unsigned char* const* in0, const unsigned char* const* in1, int w, int hh, const float* ts,
                            unsigned char* const* out) {
    GUARD_BEGIN
    if (!h) return RIFE_B200_ERR_ARG;
    return map_err(h->eng->process_batch(n, in0, in1, w, hh, ts, out));
    GUARD_END
}

int rife_b200_set_option(rife_b200_t* h, const char* key, int value) {
    GUARD_BEGIN
    if (!h || !key) return RIFE_B200_ERR_ARG;
    return map_err(h->eng->set_option(key, value));
    GUARD_END
}

int rife_b200_weights_size(rife_b200_t* h, size_t* bytes) {
    GUARD_BEGIN
    if (!h || !bytes) return RIFE_B200_ERR_ARG;
    if (h->eng->packed().empty()) return RIFE_B200_ERR_STATE;
    *bytes = h->eng->packed().size();
    return RIFE_B200_OK;
    GUARD_END
}

int rife_b200_weights_export(rife_b200_t* h, void* dst, size_t bytes) {
    GUARD_BEGIN
    if (!h || !dst) return RIFE_B200_ERR_ARG;
    const std::string& p = h->eng->packed();
    if (p.empty()) return RIFE_B200_ERR_STATE;
    if (bytes < p.size()) return RIFE_B200_ERR_ARG;
    memcpy(dst, p.data(), p.size());
    return RIFE_B200_OK;
    GUARD_END
}

int rife_b200_load_packed(rife_b200_t* h, const void* src, size_t bytes) {
    GUARD_BEGIN
    if (!h || !src) return RIFE_B200_ERR_ARG;
    return map_err(h->eng->load_packed(src, bytes));
    GUARD_END
}

unsigned long long rife_b200_launch_count(void) { return rife::g_launch_count; }

const char* rife_b200_last_error(rife_b200_t* h) { return h ? h->eng->last_error.c_str() : "null handle"; }

void rife_b200_destroy(rife_b200_t* h) {
    if (!h) return;
    try {
        delete h->eng;
        delete h;
    } catch (...) {
    }
}

}  // extern "C"
