// emu_layout.cpp -- TEST INFRASTRUCTURE.  The layout-conversion kernels between the generic executor's planar fp32 blobs and the
// tensor-core path's C8 planar split-fp16 tensors (csrc/tc_conv.cu, section "layout conversion kernels": planar_to_c8_kernel,
// c8_to_planar_kernel and their launchers) compiled for the HOST, unmodified: tests/test_layout_kernels_cpu.py cuts that section
// (and the pack2 helper) out of tc_conv.cu and rewrites its two launch statements; a launch is a loop over the grid (the kernels
// have no barriers).  No GPU, no CUDA runtime.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <cuda_fp16.h>
#include <cuda_runtime.h>  // types only

#undef __global__
#undef __device__
#undef __forceinline__
#define __global__
#define __device__
#define __forceinline__ inline

static thread_local uint3 blockIdx, threadIdx;
static thread_local dim3 blockDim, gridDim;

template <class F>
static void emu_launch(dim3 g, int nthreads, F body) {
    gridDim = g;
    blockDim = dim3((unsigned)nthreads, 1, 1);
    for (unsigned x = 0; x < g.x; x++)
        for (int t = 0; t < nthreads; t++) {
            blockIdx = make_uint3(x, 0, 0);
            threadIdx = make_uint3((unsigned)t, 0, 0);
            body();
        }
}

namespace rife {
unsigned long long g_launch_count = 0;
#include "layout_kernels_emu.inc"  // generated: pack2 + the layout section of csrc/tc_conv.cu
}  // namespace rife

extern "C" int emu_planar_to_c8(const float* in, unsigned short* out, int C, int H, int W, int split, int Cpad, int s2d) {
    rife::launch_planar_to_c8(in, reinterpret_cast<__half*>(out), C, H, W, split, nullptr, Cpad, s2d);
    return 0;
}
extern "C" int emu_c8_to_planar(const unsigned short* in, float* out, int C, int H, int W, int split, int Cpad, int s2d) {
    rife::launch_c8_to_planar(reinterpret_cast<const __half*>(in), out, C, H, W, split, nullptr, Cpad, s2d);
    return 0;
}
