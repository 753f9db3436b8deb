// ncnn_compat/benchmark.h -- included by src/main.cpp:96 and src/rife.cpp:7, unused there.
#pragma once
