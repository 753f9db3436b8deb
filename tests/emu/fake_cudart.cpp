// fake_cudart.cpp -- TEST INFRASTRUCTURE.  The subset of the CUDA runtime the engine's host code calls (csrc/engine.cu, exec.cu,
// capi.cu), with HOST semantics: device memory is host memory, copies are memcpy, streams and events are handles whose work has
// always completed (the emulated kernel launches run synchronously in the calling thread).  Linked instead of libcudart into the
// whole-engine host build of tests/test_engine_emu_cpu.py, where the engine's real control flow -- plans, arenas, lanes, the
// orientation fork / join of the TTA modes, the frame table, staging -- runs against the oracle without a GPU.  Never part of the product.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

#include <cuda_runtime.h>

extern "C" {

// exact sizes (no slack behind a buffer): under AddressSanitizer an out-of-bounds access of an emulated kernel is then a report.
// Blocks are remembered by kind so that cudaPointerGetAttributes can tell "device", pinned and pageable memory apart.
static std::mutex g_mu;
static std::map<const void*, std::pair<size_t, int>> g_blocks;  // start -> (size, 1 = device, 2 = pinned host)
static void* emu_alloc(size_t n, int kind) {
    void* p = nullptr;
    if (posix_memalign(&p, 256, n ? n : 1) != 0) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    g_blocks[p] = std::make_pair(n ? n : 1, kind);
    return p;
}
static void emu_free(void* p) {
    if (!p) return;
    { std::lock_guard<std::mutex> lk(g_mu); g_blocks.erase(p); }
    free(p);
}
cudaError_t cudaMalloc(void** p, size_t n) { *p = emu_alloc(n, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { emu_free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned int) { *p = emu_alloc(n, 2); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaMallocHost(void** p, size_t n) { return cudaHostAlloc(p, n, 0); }
cudaError_t cudaFreeHost(void* p) { emu_free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }

cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)malloc(16); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(16); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* q) {
    memset(a, 0, sizeof *a);
    a->type = cudaMemoryTypeUnregistered;  // a caller's own buffer is pageable memory: the engine's staging path is exercised
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_blocks.upper_bound(q);
    if (it != g_blocks.begin()) {
        --it;
        if ((const char*)q < (const char*)it->first + it->second.first) a->type = it->second.second == 1 ? cudaMemoryTypeDevice : cudaMemoryTypeHost;
    }
    return cudaSuccess;
}

}  // extern "C"

// cudaGetDeviceProperties is a versioned symbol behind a macro in newer toolkits: define whatever name the header maps it to
#ifdef cudaGetDeviceProperties
#undef cudaGetDeviceProperties
extern "C" cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    p->multiProcessorCount = 148;
    p->major = 10;
    return cudaSuccess;
}
#endif
extern "C" cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    p->multiProcessorCount = 148;
    p->major = 10;
    return cudaSuccess;
}
