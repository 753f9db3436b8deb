// shim_demo.cpp -- minimal caller written the way src/main.cpp uses the engine (Mat construction at
// main.cpp:187/332, RIFE ctor + load at :825-827, process at :360).  Reads two raw RGB files, writes one.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "cpu.h"
#include "gpu.h"
#include "platform.h"
#include "rife.h"

int main(int argc, char** argv)
{
    if (argc == 2 && !strcmp(argv[1], "--probe")) {
        ncnn::create_gpu_instance();
        printf("{\"gpu_count\": %d, \"cpu_count\": %d}\n", ncnn::get_gpu_count(), ncnn::get_cpu_count());
        RIFE cpu_mode(-1);  // must refuse without crashing
        ncnn::Mat a(4, 2, (size_t)3, 3), b;
        b = a;
        printf("{\"mat_ok\": %d}\n", (int)(!a.empty() && b.data == a.data && a.w == 4 && a.h == 2 && a.elempack == 3));
        ncnn::destroy_gpu_instance();
        return 0;
    }
    if (argc < 9) {
        fprintf(stderr, "usage: shim_demo modeldir v1|v2|v4 w h in0.rgb in1.rgb timestep out.rgb [tta] [tta_temporal] [uhd]\n");
        return 2;
    }
    std::string modeldir = argv[1], fam = argv[2];
    int w = atoi(argv[3]), h = atoi(argv[4]);
    float t = (float)atof(argv[7]);
    bool tta = argc > 9 && atoi(argv[9]), ttat = argc > 10 && atoi(argv[10]), uhd = argc > 11 && atoi(argv[11]);
    size_t n = (size_t)w * h * 3;
    std::vector<unsigned char> p0(n), p1(n);
    FILE* f = fopen(argv[5], "rb");
    if (!f || fread(p0.data(), 1, n, f) != n) return 2;
    fclose(f);
    f = fopen(argv[6], "rb");
    if (!f || fread(p1.data(), 1, n, f) != n) return 2;
    fclose(f);
    ncnn::create_gpu_instance();
    RIFE rife(ncnn::get_default_gpu_index(), tta, ttat, uhd, 1, fam == "v2", fam == "v4");
    if (rife.load(modeldir)) return 3;
    ncnn::Mat in0(w, h, (void*)p0.data(), (size_t)3, 3), in1(w, h, (void*)p1.data(), (size_t)3, 3);
    ncnn::Mat out(w, h, (size_t)3, 3);
    if (rife.process(in0, in1, t, out)) return 4;
    f = fopen(argv[8], "wb");
    fwrite(out.data, 1, n, f);
    fclose(f);
    ncnn::destroy_gpu_instance();
    return 0;
}
