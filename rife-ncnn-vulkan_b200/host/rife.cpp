// rife.cpp -- host shim: class RIFE (reference API) -> C ABI of librife_b200.so via dlopen.
// Error behaviour mirrors the reference where it has one (load/process return int, callers ignore it,
// src/main.cpp:360,827); gpuid == -1 (CPU mode) is refused with a message, there is no CPU fallback.
#include "rife.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

#include "gpu.h"
#include "../../include/rife_b200.h"

namespace {
struct Api {
    void* so;
    int (*device_count)(void);
    int (*create)(rife_b200_t**, int, int, int, int, int, int, int);
    int (*load)(rife_b200_t*, const char*);
    int (*load_w)(rife_b200_t*, const wchar_t*);
    int (*set_option)(rife_b200_t*, const char*, int);
    int (*process)(rife_b200_t*, const unsigned char*, const unsigned char*, int, int, float, unsigned char*);
    const char* (*last_error)(rife_b200_t*);
    void (*destroy)(rife_b200_t*);
};

Api* api() {
    static Api a;
    static bool tried = false;
    if (tried) return a.so ? &a : 0;
    tried = true;
    const char* path = getenv("RIFE_B200_LIB");
    a.so = dlopen(path ? path : "librife_b200.so", RTLD_NOW | RTLD_LOCAL);
    if (!a.so) {
        fprintf(stderr, "rife_b200: cannot load %s: %s\n", path ? path : "librife_b200.so", dlerror());
        return 0;
    }
#define BIND(field, name)                                                   \
    *(void**)(&a.field) = dlsym(a.so, name);                                \
    if (!a.field) {                                                         \
        fprintf(stderr, "rife_b200: missing symbol %s\n", name);            \
        dlclose(a.so);                                                      \
        a.so = 0;                                                           \
        return 0;                                                           \
    }
    BIND(device_count, "rife_b200_device_count")
    BIND(create, "rife_b200_create")
    BIND(load, "rife_b200_load")
    BIND(load_w, "rife_b200_load_w")
    BIND(set_option, "rife_b200_set_option")
    BIND(process, "rife_b200_process")
    BIND(last_error, "rife_b200_last_error")
    BIND(destroy, "rife_b200_destroy")
#undef BIND
    return &a;
}
}  // namespace

namespace ncnn {
int create_gpu_instance() { return api() ? 0 : -1; }
void destroy_gpu_instance() {}
int get_gpu_count() { Api* a = api(); return a ? a->device_count() : 0; }
int get_default_gpu_index() { return 0; }
}  // namespace ncnn

RIFE::RIFE(int gpuid, bool tta_mode, bool tta_temporal_mode, bool uhd_mode, int num_threads, bool rife_v2, bool rife_v4) : handle(0), create_status(-1)
{
    Api* a = api();
    if (!a) return;
    if (gpuid < 0) {
        fprintf(stderr, "rife_b200: gpuid %d requested; this engine has no CPU path (use -g 0..%d)\n", gpuid, a->device_count() - 1);
        return;
    }
    create_status = a->create(&handle, gpuid, tta_mode, tta_temporal_mode, uhd_mode, num_threads, rife_v2, rife_v4);
    if (create_status) fprintf(stderr, "rife_b200_create(gpu %d) failed: %d\n", gpuid, create_status);
#if _WIN32
    if (!create_status) a->set_option(handle, "bgr", 1);  // the Windows codecs of src/main.cpp produce / consume B,G,R
#endif
}

RIFE::~RIFE()
{
    if (handle) api()->destroy(handle);
}

int RIFE::load(const std::string& modeldir)
{
    if (!handle) return -1;
    int r = api()->load(handle, modeldir.c_str());
    if (r) fprintf(stderr, "rife_b200_load(%s) failed: %s\n", modeldir.c_str(), api()->last_error(handle));
    return r;
}

int RIFE::load(const std::wstring& modeldir)
{
    if (!handle) return -1;
    int r = api()->load_w(handle, modeldir.c_str());
    if (r) fprintf(stderr, "rife_b200_load_w failed: %s\n", api()->last_error(handle));
    return r;
}

int RIFE::set_bgr(bool bgr)
{
    if (!handle) return -1;
    return api()->set_option(handle, "bgr", bgr ? 1 : 0);
}

int RIFE::process(const ncnn::Mat& in0image, const ncnn::Mat& in1image, float timestep, ncnn::Mat& outimage) const
{
    if (!handle) return -1;
    if (timestep == 0.f) { outimage = in0image; return 0; }
    if (timestep == 1.f) { outimage = in1image; return 0; }
    if (in0image.empty() || in1image.empty() || outimage.empty() || in0image.w != in1image.w || in0image.h != in1image.h ||
        outimage.w != in0image.w || outimage.h != in0image.h || in0image.elemsize != 3 || outimage.elemsize != 3) return -1;
    int r = api()->process(handle, (const unsigned char*)in0image.data, (const unsigned char*)in1image.data, in0image.w, in0image.h, timestep,
                           (unsigned char*)outimage.data);
    if (r) fprintf(stderr, "rife_b200_process failed: %s\n", api()->last_error(handle));
    return r;
}
