// cuda_host_shim.h -- TEST INFRASTRUCTURE.  The part of the CUDA programming model the repo's plain (non-tcgen05) kernels use, for
// a HOST build of the unmodified kernel sources: a block's threads are fibers of the launching host thread, __syncthreads() is a
// barrier of the block, __shfl_xor_sync() an exchange through a per-warp barrier, `__shared__` arrays are statics (blocks run one
// after another), dynamic shared memory is a global buffer.  A launch statement `k<<<grid, block, smem, st>>>(args)` is rewritten by the tests into
// emu_launch(dim3(grid), block, smem, [&] { k(args); }).  Used by tests/emu/emu_generic.cpp, emu_hbm.cpp and the whole-engine host
// build (tests/test_engine_emu_cpu.py).  Never part of the product.
#pragma once
#include <math.h>
#include <ucontext.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
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

// ---- a block = nthreads FIBERS of the launching host thread (ucontext), scheduled round-robin --------------------------------
// __syncthreads(): the fiber yields until every fiber of the block that has not returned arrived (generation counter); the warp
// barrier behind __shfl_xor_sync works the same over the 32 lanes of a warp.  (A first version used one OS thread per CUDA
// thread and pthread barriers: correct, but a 256-thread barrier costs about a millisecond.)
// Context switch: on x86-64 six callee-saved registers and the stack pointer (swapcontext costs two sigprocmask system calls per
// switch, and a barrier of 256 fibers is about a thousand switches); elsewhere ucontext.
#if defined(__x86_64__) && !defined(EMU_NO_FAST_SWITCH)
#define EMU_FAST_SWITCH 1
__attribute__((naked, noinline)) static void emu_swap(void** /*save_sp: rdi*/, void* /*load_sp: rsi*/) {
    asm volatile(
        "pushq %rbp\n pushq %rbx\n pushq %r12\n pushq %r13\n pushq %r14\n pushq %r15\n"
        "movq %rsp, (%rdi)\n"
        "movq %rsi, %rsp\n"
        "popq %r15\n popq %r14\n popq %r13\n popq %r12\n popq %rbx\n popq %rbp\n"
        "ret\n");
}
typedef void* EmuCtx;
#else
typedef ucontext_t EmuCtx;
#endif
struct EmuSched {
    EmuCtx main_ctx;
    std::vector<EmuCtx> ctx;
    std::vector<char> stacks;
    std::vector<char> done;
    int nthreads = 0, cur = 0, ndone = 0;
    int block_arrived = 0;
    unsigned block_gen = 0;
    int warp_arrived[32], warp_done[32];
    unsigned warp_gen[32];
    float shfl[1024];
    void (*entry)(void*) = nullptr;
    void* entry_arg = nullptr;
    std::vector<unsigned char> dyn_smem;
};
static thread_local EmuSched* emu_sched = nullptr;
static unsigned char* emu_dyn_smem = nullptr;  // per translation unit; set by emu_launch before the kernel body runs
inline std::mutex& emu_launch_mutex() {        // one launch at a time: the statics that stand for `__shared__` arrays are per kernel
    static std::mutex m;
    return m;
}

static inline void emu_switch(EmuCtx* from, EmuCtx* to) {
#ifdef EMU_FAST_SWITCH
    emu_swap(from, *to);
#else
    swapcontext(from, to);
#endif
}
static inline void emu_yield() {
    EmuSched& S = *emu_sched;
    emu_switch(&S.ctx[S.cur], &S.main_ctx);
}
static inline void __syncthreads() {
    EmuSched& S = *emu_sched;
    const unsigned my = S.block_gen;
    if (++S.block_arrived + S.ndone == S.nthreads) { S.block_arrived = 0; S.block_gen++; }
    while (S.block_gen == my) emu_yield();
}
static inline void emu_warp_barrier(int w) {
    EmuSched& S = *emu_sched;
    const unsigned my = S.warp_gen[w];
    const int lanes = S.nthreads - 32 * w < 32 ? S.nthreads - 32 * w : 32;
    if (++S.warp_arrived[w] + S.warp_done[w] == lanes) { S.warp_arrived[w] = 0; S.warp_gen[w]++; }
    while (S.warp_gen[w] == my) emu_yield();
}
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {  // all lanes of the calling warp that are still running take part
    EmuSched& S = *emu_sched;
    const unsigned t = threadIdx.x;
    S.shfl[t] = v;
    emu_warp_barrier((int)(t >> 5));
    const float r = S.shfl[t ^ (unsigned)lane_mask];
    emu_warp_barrier((int)(t >> 5));
    return r;
}
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }

static void emu_fiber_main() {
    EmuSched& S = *emu_sched;
    S.entry(S.entry_arg);
    const int t = S.cur, w = t >> 5;
    S.done[t] = 1;
    S.ndone++;
    S.warp_done[w]++;
    // a returned thread no longer counts at the barriers: release the ones that were only waiting for it
    if (S.block_arrived && S.block_arrived + S.ndone == S.nthreads) { S.block_arrived = 0; S.block_gen++; }
    const int lanes = S.nthreads - 32 * w < 32 ? S.nthreads - 32 * w : 32;
    if (S.warp_arrived[w] && S.warp_arrived[w] + S.warp_done[w] == lanes) { S.warp_arrived[w] = 0; S.warp_gen[w]++; }
    emu_switch(&S.ctx[t], &S.main_ctx);  // never resumed
    abort();
}

template <class F>
static void emu_launch(dim3 g, int nthreads, size_t smem_bytes, F body) {
#ifdef EMU_SKIP_KERNELS  // host-logic-only builds (ThreadSanitizer over the engine's threading: no fibers, results are not looked at)
    (void)g; (void)nthreads; (void)smem_bytes; (void)body;
    return;
#endif
    std::lock_guard<std::mutex> lk(emu_launch_mutex());
    static thread_local EmuSched sched;
    EmuSched& S = sched;
    emu_sched = &S;
    constexpr size_t kStack = 256 * 1024;
    S.nthreads = nthreads;
    S.ctx.resize((size_t)nthreads);
    S.done.resize((size_t)nthreads);
    if (S.stacks.size() < kStack * (size_t)nthreads) S.stacks.resize(kStack * (size_t)nthreads);
    S.dyn_smem.assign(smem_bytes + 64, 0);
    emu_dyn_smem = S.dyn_smem.data();
    S.entry = [](void* p) { (*static_cast<F*>(p))(); };
    S.entry_arg = &body;
    gridDim = g;
    blockDim = dim3((unsigned)nthreads, 1, 1);
    for (unsigned z = 0; z < g.z; z++)
        for (unsigned y = 0; y < g.y; y++)
            for (unsigned x = 0; x < g.x; x++) {
                blockIdx = make_uint3(x, y, z);
                S.ndone = 0;
                S.block_arrived = 0;
                for (int wv = 0; wv < 32; wv++) S.warp_arrived[wv] = S.warp_done[wv] = 0;
                for (int t = 0; t < nthreads; t++) {
                    S.done[t] = 0;
#ifdef EMU_FAST_SWITCH
                    // a fresh fiber: six zeroed callee-saved registers below the entry address; after the `ret` of emu_swap the stack
                    // pointer is 8 below a 16-byte boundary, as at any function entry
                    uintptr_t top = ((uintptr_t)(S.stacks.data() + kStack * (size_t)(t + 1))) & ~(uintptr_t)15;
                    void** sp = (void**)top;
                    *--sp = nullptr;                      // where a return address would sit (emu_fiber_main never returns)
                    *--sp = (void*)&emu_fiber_main;       // popped by `ret`
                    for (int r = 0; r < 6; r++) *--sp = nullptr;
                    S.ctx[t] = (void*)sp;
#else
                    getcontext(&S.ctx[t]);
                    S.ctx[t].uc_stack.ss_sp = S.stacks.data() + kStack * (size_t)t;
                    S.ctx[t].uc_stack.ss_size = kStack;
                    S.ctx[t].uc_link = nullptr;
                    makecontext(&S.ctx[t], emu_fiber_main, 0);
#endif
                }
                while (S.ndone < nthreads)
                    for (int t = 0; t < nthreads; t++) {
                        if (S.done[t]) continue;
                        S.cur = t;
                        threadIdx = make_uint3((unsigned)t, 0, 0);
                        emu_switch(&S.main_ctx, &S.ctx[t]);
                    }
            }
    emu_sched = nullptr;
}
template <class F>
static void emu_launch(dim3 g, int nthreads, F body) { emu_launch(g, nthreads, 0, body); }
