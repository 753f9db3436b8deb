// ncnn_compat/cpu.h -- get_cpu_count() as used at /root/reference/src/main.cpp:788
#pragma once
#include <unistd.h>
namespace ncnn {
inline int get_cpu_count() {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
}  // namespace ncnn
