// emu_engine_stubs.cpp -- TEST INFRASTRUCTURE, part of the whole-engine host build (tests/test_engine_emu_cpu.py).
// What cannot run on a host: the tcgen05 convolution (csrc/tc_conv.cu's kernel and launcher) and the hand-scheduled runner built
// on it (csrc/fused_v46.cu).  Their entry points exist here and REFUSE, so the engine takes the paths it takes whenever a layer or
// a model is not tensor-core eligible: fp32 kernels (precision tier 0) under the generic executor.  The host-side parts of
// tc_conv.cu -- weight packers, layout kernels -- are the real ones (cut out of the source by the test).
#include "cuda_host_shim.h"

#include "fused_v46.h"
#include "kernels.h"
#include "rife_b200.h"
#include "tc_conv.h"

// The host build has no tensor cores, so a handle starts at precision tier 0 (fp32 kernels) instead of tier 1: capi.cu is compiled
// with -Drife_b200_create=rife_b200_create_tier1 and this is the exported constructor.  (A caller that cannot set options -- the
// reference's CLI -- then works against the host build too.)
extern "C" int rife_b200_create_tier1(rife_b200_t** handle, int gpuid, int tta_mode, int tta_temporal_mode, int uhd_mode, int num_threads, int rife_v2, int rife_v4);
extern "C" int rife_b200_create(rife_b200_t** handle, int gpuid, int tta_mode, int tta_temporal_mode, int uhd_mode, int num_threads, int rife_v2, int rife_v4) {
    const int r = rife_b200_create_tier1(handle, gpuid, tta_mode, tta_temporal_mode, uhd_mode, num_threads, rife_v2, rife_v4);
    if (r == 0) rife_b200_set_option(*handle, "precision", 0);
    return r;
}

namespace rife {

int launch_tc_conv(TcConvArgs, const void*, cudaStream_t) { return -100; }  // no tcgen05 on the host

V46Runner::~V46Runner() {}
int V46Runner::init(const Net*, const NetRunner*, std::string& err) { err = "the fused tcgen05 path does not exist in the host build"; return -1; }
void V46Runner::set_ktime(int) {}
std::string V46Runner::stage_report() const { return "batches\t0\n"; }
int V46Runner::run(const uint8_t*, const uint8_t*, int, int, float, uint8_t*, cudaStream_t, std::string& err) { err = "host build"; return -1; }
int V46Runner::run_batch(int, const uint8_t* const*, const uint8_t* const*, int, int, const float*, uint8_t* const*, cudaStream_t, std::string& err) { err = "host build"; return -1; }

#include "tc_host_section.inc"  // generated: pack2, tc_conv_tile_rows and everything of csrc/tc_conv.cu from its layout kernels on
                                // (that text closes namespace rife itself)
