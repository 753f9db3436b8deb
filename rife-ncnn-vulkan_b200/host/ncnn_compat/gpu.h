// ncnn_compat/gpu.h -- device enumeration entry points src/main.cpp calls (:774-799, :914), answered by the
// CUDA library through rife_b200_device_count().
#pragma once
namespace ncnn {
int create_gpu_instance();
void destroy_gpu_instance();
int get_gpu_count();
int get_default_gpu_index();
}  // namespace ncnn
