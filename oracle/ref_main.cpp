// oracle/ref_main.cpp -- driver for oracle/_ref/ref_rife_<isa>: the reference's own `-g -1` CPU path.
//
// TEST INFRASTRUCTURE ONLY.  Links the reference's vendored ncnn (compiled in place by oracle/build_ref.py)
// and *includes* the CPU function bodies extracted at build time from /root/reference/src/rife.cpp
// (RIFE::process_cpu :1214-2460, RIFE::process_v4_cpu :3204-4401) and /root/reference/src/warp.cpp
// (Warp::forward CPU overload :96-168).  Only the Vulkan-free scaffolding around them is written here:
//   * class Warp without the Vulkan members of rife_ops.h:12-25,
//   * class RIFE with the data members of rife.h:31-51 that the CPU functions touch,
//   * RIFE::load restating the vkdev==0 branch of rife.cpp:127-379 (Option block :130-136, custom layer
//     registration :146-148, uhd helper layers :294-332, v2 Slice helper :334-351).
// The reference itself cannot be compiled unmodified here: rife.h/rife_ops.h need ncnn's Vulkan types and
// src/CMakeLists.txt:19 requires the Vulkan SDK, neither of which exists in this image.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>

#include "net.h"
#include "layer.h"
#include "cpu.h"

using namespace ncnn;

class Warp : public ncnn::Layer
{
public:
    Warp() { one_blob_only = false; support_inplace = false; }
    virtual int forward(const std::vector<ncnn::Mat>& bottom_blobs, std::vector<ncnn::Mat>& top_blobs, const ncnn::Option& opt) const;
};

#include "warp_cpu_extract.inc"

DEFINE_LAYER_CREATOR(Warp)

class RIFE
{
public:
    RIFE(bool _tta, bool _tta_temporal, bool _uhd, int _num_threads, bool _v2, bool _v4)
        : rife_uhd_downscale_image(0), rife_uhd_upscale_flow(0), rife_uhd_double_flow(0), rife_v2_slice_flow(0),
          tta_mode(_tta), tta_temporal_mode(_tta_temporal), uhd_mode(_uhd), num_threads(_num_threads), rife_v2(_v2), rife_v4(_v4) {}
    int load(const std::string& modeldir);
    int process(const ncnn::Mat& a, const ncnn::Mat& b, float t, ncnn::Mat& o) const
    {
        return rife_v4 ? process_v4_cpu(a, b, t, o) : process_cpu(a, b, t, o); // rife.cpp:383-390
    }
    int process_cpu(const ncnn::Mat& in0image, const ncnn::Mat& in1image, float timestep, ncnn::Mat& outimage) const;
    int process_v4_cpu(const ncnn::Mat& in0image, const ncnn::Mat& in1image, float timestep, ncnn::Mat& outimage) const;

    ncnn::Net flownet;
    ncnn::Net contextnet;
    ncnn::Net fusionnet;
    ncnn::Layer* rife_uhd_downscale_image;
    ncnn::Layer* rife_uhd_upscale_flow;
    ncnn::Layer* rife_uhd_double_flow;
    ncnn::Layer* rife_v2_slice_flow;
    bool tta_mode;
    bool tta_temporal_mode;
    bool uhd_mode;
    int num_threads;
    bool rife_v2;
    bool rife_v4;
};

static int load_net(ncnn::Net& net, const std::string& dir, const char* name)
{
    std::string p = dir + "/" + name + ".param", b = dir + "/" + name + ".bin";
    if (net.load_param(p.c_str())) return -1;
    if (net.load_model(b.c_str())) return -1;
    return 0;
}

static ncnn::Layer* make_layer(const char* type, const ncnn::ParamDict& pd, const ncnn::Option& opt)
{
    ncnn::Layer* l = ncnn::create_layer(type);
    l->load_param(pd);
    l->create_pipeline(opt);
    return l;
}

int RIFE::load(const std::string& modeldir)
{
    ncnn::Option opt; // vkdev == 0 values of rife.cpp:130-136
    opt.num_threads = num_threads;
    opt.use_vulkan_compute = false;
    opt.use_fp16_packed = false;
    opt.use_fp16_storage = false;
    opt.use_fp16_arithmetic = false;
    opt.use_int8_storage = true;
    flownet.opt = opt;
    contextnet.opt = opt;
    fusionnet.opt = opt;
    flownet.register_custom_layer("rife.Warp", Warp_layer_creator);
    contextnet.register_custom_layer("rife.Warp", Warp_layer_creator);
    fusionnet.register_custom_layer("rife.Warp", Warp_layer_creator);
    int ret = load_net(flownet, modeldir, "flownet");
    if (!rife_v4)
    {
        ret |= load_net(contextnet, modeldir, "contextnet");
        ret |= load_net(fusionnet, modeldir, "fusionnet");
    }
    if (uhd_mode)
    {
        { ncnn::ParamDict pd; pd.set(0, 2); pd.set(1, 0.5f); pd.set(2, 0.5f); rife_uhd_downscale_image = make_layer("Interp", pd, opt); }
        { ncnn::ParamDict pd; pd.set(0, 2); pd.set(1, 2.f); pd.set(2, 2.f); rife_uhd_upscale_flow = make_layer("Interp", pd, opt); }
        { ncnn::ParamDict pd; pd.set(0, 2); pd.set(1, 1); pd.set(2, 2.f); rife_uhd_double_flow = make_layer("BinaryOp", pd, opt); }
    }
    if (rife_v2)
    {
        ncnn::Mat slice_points(2);
        slice_points.fill<int>(-233);
        ncnn::ParamDict pd;
        pd.set(0, slice_points);
        pd.set(1, 0);
        rife_v2_slice_flow = make_layer("Slice", pd, opt);
    }
    return ret;
}

#include "rife_cpu_extract.inc"

static std::vector<unsigned char> read_file(const char* path, size_t expect)
{
    std::vector<unsigned char> v(expect);
    FILE* fp = fopen(path, "rb");
    if (!fp || fread(v.data(), 1, expect, fp) != expect)
    {
        fprintf(stderr, "ref_rife: cannot read %zu bytes from %s\n", expect, path);
        exit(2);
    }
    fclose(fp);
    return v;
}

static void write_file(const char* path, const void* p, size_t n)
{
    FILE* fp = fopen(path, "wb");
    if (!fp || fwrite(p, 1, n, fp) != n)
    {
        fprintf(stderr, "ref_rife: cannot write %s\n", path);
        exit(2);
    }
    fclose(fp);
}

// --extract: plain (non-TTA) v4 extractor run dumping named intermediate blobs as dense float32 CHW,
// preproc as rife.cpp:4152-4214.  Debug aid for layer-wise comparisons.
static int extract_blobs(const RIFE& r, const unsigned char* p0, const unsigned char* p1, int w, int h, float t,
                         const std::string& names, const std::string& prefix)
{
    int wp = (w + 31) / 32 * 32, hp = (h + 31) / 32 * 32;
    ncnn::Mat in[2];
    const unsigned char* px[2] = {p0, p1};
    for (int k = 0; k < 2; k++)
    {
        ncnn::Mat f = ncnn::Mat::from_pixels(px[k], ncnn::Mat::PIXEL_RGB, w, h);
        in[k].create(wp, hp, 3);
        in[k].fill(0.f);
        for (int q = 0; q < 3; q++)
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++)
                    in[k].channel(q).row(y)[x] = f.channel(q).row(y)[x] * (1 / 255.f);
    }
    ncnn::Mat ts(wp, hp, 1);
    ts.fill(t);
    ncnn::Extractor ex = r.flownet.create_extractor();
    ex.input("in0", in[0]);
    ex.input("in1", in[1]);
    ex.input("in2", ts);
    size_t pos = 0;
    while (pos < names.size())
    {
        size_t e = names.find(',', pos);
        if (e == std::string::npos) e = names.size();
        std::string n = names.substr(pos, e - pos);
        pos = e + 1;
        ncnn::Mat m;
        if (ex.extract(n.c_str(), m)) { fprintf(stderr, "extract %s failed\n", n.c_str()); return 3; }
        std::vector<float> dense((size_t)m.w * m.h * m.c);
        for (int q = 0; q < m.c; q++)
            for (int y = 0; y < m.h; y++)
                memcpy(&dense[((size_t)q * m.h + y) * m.w], m.channel(q).row(y), sizeof(float) * m.w);
        std::string fn = prefix + n + ".f32";
        write_file(fn.c_str(), dense.data(), dense.size() * 4);
        printf("{\"blob\": \"%s\", \"c\": %d, \"h\": %d, \"w\": %d}\n", n.c_str(), m.c, m.h, m.w);
    }
    return 0;
}

int main(int argc, char** argv)
{
    std::string model, family = "v4", in0, in1, out, extract, prefix = "blob_";
    int w = 0, h = 0, threads = ncnn::get_cpu_count(), repeat = 1, warmup = 0;
    bool tta = false, ttat = false, uhd = false;
    float t = 0.5f;
    for (int i = 1; i < argc; i++)
    {
        std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--model") model = next();
        else if (a == "--family") family = next();
        else if (a == "--in0") in0 = next();
        else if (a == "--in1") in1 = next();
        else if (a == "--out") out = next();
        else if (a == "--w") w = atoi(next());
        else if (a == "--h") h = atoi(next());
        else if (a == "--t") t = (float)atof(next());
        else if (a == "--threads") threads = atoi(next());
        else if (a == "--repeat") repeat = atoi(next());
        else if (a == "--warmup") warmup = atoi(next());
        else if (a == "--tta") tta = true;
        else if (a == "--tta-temporal") ttat = true;
        else if (a == "--uhd") uhd = true;
        else if (a == "--extract") extract = next();
        else if (a == "--prefix") prefix = next();
        else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
    }
    if (model.empty() || in0.empty() || in1.empty() || w <= 0 || h <= 0)
    {
        fprintf(stderr, "usage: ref_rife --model DIR --family v1|v2|v4 --w W --h H --in0 a.rgb --in1 b.rgb [--out o.rgb] [--t 0.5]\n"
                        "       [--tta] [--tta-temporal] [--uhd] [--threads N] [--warmup W] [--repeat R] [--extract blob,blob --prefix p]\n");
        return 2;
    }
    bool v2 = family == "v2", v4 = family == "v4";
    RIFE rife(tta, ttat, uhd, threads, v2, v4);
    if (rife.load(model)) { fprintf(stderr, "ref_rife: load(%s) failed\n", model.c_str()); return 2; }
    std::vector<unsigned char> a = read_file(in0.c_str(), (size_t)w * h * 3), b = read_file(in1.c_str(), (size_t)w * h * 3);
    if (!extract.empty())
        return extract_blobs(rife, a.data(), b.data(), w, h, t, extract, prefix);
    std::vector<unsigned char> o((size_t)w * h * 3);
    ncnn::Mat m0(w, h, (void*)a.data(), (size_t)3, 3), m1(w, h, (void*)b.data(), (size_t)3, 3); // main.cpp:187
    ncnn::Mat mo(w, h, (void*)o.data(), (size_t)3, 3);
    std::string times;
    for (int r = 0; r < warmup + repeat; r++)
    {
        auto t0 = std::chrono::steady_clock::now();
        rife.process(m0, m1, t, mo);
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (r >= warmup) { char buf[32]; snprintf(buf, sizeof buf, "%s%.6f", times.empty() ? "" : ", ", s); times += buf; }
    }
    // t == 0 / 1 rebinds the Mat to an input (rife.cpp:3206-3216): honour that when writing
    if (!out.empty()) write_file(out.c_str(), mo.data, (size_t)w * h * 3);
    printf("{\"sec_per_frame\": [%s], \"threads\": %d, \"w\": %d, \"h\": %d}\n", times.c_str(), threads, w, h);
    return 0;
}
