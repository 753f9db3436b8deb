// ncnn_compat/net.h -- rife.h includes "net.h" for ncnn::Mat only (this engine has no ncnn::Net).
#pragma once
#include "mat.h"
