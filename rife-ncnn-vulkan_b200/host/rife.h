// rife.h -- drop-in replacement for /root/reference/src/rife.h:11-52: same class name, constructor and
// load/process signatures, so the reference's src/main.cpp compiles and runs against it unchanged.
// Everything behind it is librife_b200.so (include/rife_b200.h), bound lazily with dlopen.
#ifndef RIFE_H
#define RIFE_H

#include <string>

#include "net.h"  // ncnn_compat: ncnn::Mat

struct rife_b200;

class RIFE
{
public:
    RIFE(int gpuid, bool tta_mode = false, bool tta_temporal_mode = false, bool uhd_mode = false, int num_threads = 1, bool rife_v2 = false, bool rife_v4 = false);
    ~RIFE();

    int load(const std::string& modeldir);
    // the reference's Windows build declares load(const std::wstring&) instead (src/rife.h:21-25) and hands over / expects
    // B,G,R frames (src/rife.cpp:438-444, rife_preproc.comp:13,53-56): both are available here on every platform, the BGR
    // channel order is switched on automatically under _WIN32 (or with set_bgr)
    int load(const std::wstring& modeldir);
    int set_bgr(bool bgr);

    // in0image / in1image: w x h packed RGB u8 (elemsize 3, elempack 3); outimage: caller-allocated same shape.
    // timestep 0 / 1 rebinds outimage to the input Mat, as the reference does (src/rife.cpp:3206-3216).
    int process(const ncnn::Mat& in0image, const ncnn::Mat& in1image, float timestep, ncnn::Mat& outimage) const;

private:
    RIFE(const RIFE&);
    RIFE& operator=(const RIFE&);
    rife_b200* handle;
    int create_status;
};

#endif // RIFE_H
