// cuda_host_shim.h -- TEST INFRASTRUCTURE.  The part of the CUDA programming model the repo's plain (non-tcgen05) kernels use, for
// a HOST build of the unmodified kernel sources: a block's threads are real host threads, __syncthreads() is a barrier of the
// block, __shfl_xor_sync() an exchange through a per-warp barrier, `__shared__` arrays are statics (blocks run one after another),
// dynamic shared memory is a global buffer.  A launch statement `k<<<grid, block, smem, st>>>(args)` is rewritten by the tests into
// emu_launch(dim3(grid), block, smem, [&] { k(args); }).  Used by tests/emu/emu_generic.cpp, emu_hbm.cpp and the whole-engine host
// build (tests/test_engine_emu_cpu.py).  Never part of the product.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <thread>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>  // types and prototypes only

#undef __global__
#undef __device__
#undef __shared__
#undef __launch_bounds__
#undef __forceinline__
#define __global__
#define __device__
#define __shared__ static
#define __launch_bounds__(...)
#define __forceinline__ inline
#ifndef EMU_KEEP_CUDA_CALLS  // kernel-only builds: the two runtime calls the conv launcher makes
#define cudaFuncSetAttribute(...) cudaSuccess
#define cudaGetDevice(p) (*(p) = 0, cudaSuccess)
#endif

static thread_local uint3 blockIdx, threadIdx;
static thread_local dim3 blockDim, gridDim;

struct EmuState {
    pthread_barrier_t block_barrier;
    pthread_barrier_t warp_barrier[32];
    float shfl[1024];
    std::vector<unsigned char> dyn_smem;
    std::mutex launch_mu;  // one launch at a time (the statics that stand for shared memory are per kernel, not per launch)
};
inline EmuState& emu_state() {
    static EmuState s;
    return s;
}
static unsigned char* emu_dyn_smem = nullptr;  // per translation unit; set by emu_launch before the kernel body runs

static inline void __syncthreads() { pthread_barrier_wait(&emu_state().block_barrier); }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {  // all 32 lanes of the calling warp take part (as on the device)
    EmuState& e = emu_state();
    const unsigned t = threadIdx.x;
    e.shfl[t] = v;
    pthread_barrier_wait(&e.warp_barrier[t >> 5]);
    const float r = e.shfl[t ^ (unsigned)lane_mask];
    pthread_barrier_wait(&e.warp_barrier[t >> 5]);
    return r;
}
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }

template <class F>
static void emu_launch(dim3 g, int nthreads, size_t smem_bytes, F body) {
    EmuState& e = emu_state();
    std::lock_guard<std::mutex> lk(e.launch_mu);
    e.dyn_smem.assign(smem_bytes + 64, 0);
    emu_dyn_smem = e.dyn_smem.data();
    pthread_barrier_init(&e.block_barrier, nullptr, (unsigned)nthreads);
    for (int wv = 0; wv < (nthreads + 31) / 32; wv++) pthread_barrier_init(&e.warp_barrier[wv], nullptr, 32);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([=, &e]() {
            gridDim = g;
            blockDim = dim3((unsigned)nthreads, 1, 1);
            threadIdx = make_uint3((unsigned)t, 0, 0);
            for (unsigned z = 0; z < g.z; z++)
                for (unsigned y = 0; y < g.y; y++)
                    for (unsigned x = 0; x < g.x; x++) {
                        blockIdx = make_uint3(x, y, z);
                        body();
                        pthread_barrier_wait(&e.block_barrier);  // the next block reuses the shared-memory statics
                    }
        });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&e.block_barrier);
    for (int wv = 0; wv < (nthreads + 31) / 32; wv++) pthread_barrier_destroy(&e.warp_barrier[wv]);
}
template <class F>
static void emu_launch(dim3 g, int nthreads, F body) { emu_launch(g, nthreads, 0, body); }
