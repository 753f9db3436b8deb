#!/usr/bin/env python3
"""Build oracle/_ref/: the reference's own `-g -1` CPU path, compiled from the sources where they lie.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Nothing under rife-ncnn-vulkan_b200/ may call this.

What this does (no cmake, no reference build system; the recipe is this script):
  * writes the handful of configuration headers ncnn's build would have generated (platform.h,
    ncnn_export.h, layer_declaration.h, layer_registry.h, layer_type_enum.h) into oracle/_ref/gen/ --
    these are *our* configuration of the library (CPU only, no Vulkan, no runtime dispatch), listing only
    the layer types the RIFE models use;
  * extracts, at build time, the CPU-only function bodies `RIFE::process_cpu`, `RIFE::process_v4_cpu`
    (/root/reference/src/rife.cpp:1214-2460, 3204-4401) and `Warp::forward` (CPU overload,
    /root/reference/src/warp.cpp:96-168) into oracle/_ref/gen/*.inc (git-ignored, never committed), so the
    executable runs the reference's *own* orchestration text, not a paraphrase of it;
  * compiles ncnn's core + the needed generic and x86 layer sources in place with g++ (-Ofast -ffast-math
    -fopenmp, the flags ncnn's Release build uses, src/ncnn/src/CMakeLists.txt:302-307) once per ISA
    variant (avx2 / avx512) and links `oracle/_ref/ref_rife_<isa>` with oracle/ref_main.cpp;
  * copies the model directories the tests/bench need into oracle/_ref/models/ (data, git-ignored; travels
    to the GPU box with the snapshot like the binaries do).

Usage: python oracle/build_ref.py [--ref /root/reference] [--isa avx2,avx512] [--jobs N] [--models a,b,c]
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

# layer types present in the 11 RIFE model dirs + the ones ncnn creates internally for them
LAYERS = ["Input", "Convolution", "Deconvolution", "PReLU", "ReLU", "Interp", "PixelShuffle", "Crop", "Concat",
          "Split", "BinaryOp", "Eltwise", "Sigmoid", "Clip", "Pooling", "InnerProduct", "UnaryOp", "Slice",
          "Padding", "Packing", "Cast", "Flatten", "Noop"]

CORE = ["allocator.cpp", "benchmark.cpp", "blob.cpp", "cpu.cpp", "datareader.cpp", "layer.cpp", "mat.cpp",
        "mat_pixel.cpp", "mat_pixel_resize.cpp", "modelbin.cpp", "net.cpp", "option.cpp", "paramdict.cpp"]

ISA_FLAGS = {
    "avx2": ["-mavx2", "-mfma", "-mf16c"],
    "avx512": ["-mavx512f", "-mavx512cd", "-mavx512bw", "-mavx512dq", "-mavx512vl", "-mfma", "-mf16c", "-mavx2"],
}

PLATFORM_ON = {"NCNN_STDIO", "NCNN_STRING", "NCNN_THREADS", "NCNN_PIXEL", "NCNN_PLATFORM_API", "NCNN_FORCE_INLINE",
               "NCNN_AVX", "NCNN_FMA", "NCNN_F16C", "NCNN_AVX2", "NCNN_BF16"}


def sh(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + "\n")
        raise SystemExit("build_ref: command failed")
    return r.stdout


def all_layer_classes(ncnn_src):
    txt = open(os.path.join(ncnn_src, "CMakeLists.txt")).read()
    return re.findall(r"^ncnn_add_layer\((\w+)", txt, flags=re.M)


def gen_headers(ncnn_src, gen, isa):
    os.makedirs(gen, exist_ok=True)
    on = set(PLATFORM_ON)
    if isa == "avx512":
        on.add("NCNN_AVX512")
    tmpl = open(os.path.join(ncnn_src, "platform.h.in")).read()
    tmpl = re.sub(r"#cmakedefine01 (\w+)", lambda m: "#define %s %d" % (m.group(1), 1 if m.group(1) in on else 0), tmpl)
    tmpl = re.sub(r"#cmakedefine NCNN_VERSION_STRING.*", '#define NCNN_VERSION_STRING "oracle-ref"', tmpl)
    open(os.path.join(gen, "platform.h"), "w").write(tmpl)
    open(os.path.join(gen, "ncnn_export.h"), "w").write(
        "#ifndef NCNN_EXPORT_H\n#define NCNN_EXPORT_H\n#define NCNN_EXPORT\n#define NCNN_NO_EXPORT\n"
        "#define NCNN_DEPRECATED\n#endif\n")
    classes = all_layer_classes(ncnn_src)
    decl, reg, enum = [], [], []
    for i, cls in enumerate(classes):
        name = cls.lower()
        enum.append("%s = %d," % (cls, i))
        if cls not in LAYERS:
            reg.append('{"%s", 0},' % cls)
            continue
        has_x86 = os.path.exists(os.path.join(ncnn_src, "layer", "x86", name + "_x86.cpp"))
        inc = ['#include "layer/%s.h"' % name]
        bases = ["virtual public %s" % cls]
        cp = ["        { int ret = %s::create_pipeline(opt); if (ret) return ret; }" % cls]
        dp = ["        { int ret = %s::destroy_pipeline(opt); if (ret) return ret; }" % cls]
        if has_x86:
            inc.append('#include "layer/x86/%s_x86.h"' % name)
            bases.append("virtual public %s_x86" % cls)
            cp.append("        { int ret = %s_x86::create_pipeline(opt); if (ret) return ret; }" % cls)
            dp.insert(0, "        { int ret = %s_x86::destroy_pipeline(opt); if (ret) return ret; }" % cls)
        decl.append("\n".join(inc) + "\nnamespace ncnn {\nclass %s_final : %s\n{\npublic:\n"
                    "    virtual int create_pipeline(const Option& opt) {\n%s\n        return 0;\n    }\n"
                    "    virtual int destroy_pipeline(const Option& opt) {\n%s\n        return 0;\n    }\n};\n"
                    "DEFINE_LAYER_CREATOR(%s_final)\n} // namespace ncnn\n"
                    % (cls, ", ".join(bases), "\n".join(cp), "\n".join(dp), cls))
        reg.append('{"%s", %s_final_layer_creator},' % (cls, cls))
    open(os.path.join(gen, "layer_declaration.h"), "w").write("\n".join(decl))
    open(os.path.join(gen, "layer_registry.h"), "w").write(
        "static const layer_registry_entry layer_registry[] = {\n" + "\n".join(reg) + "\n};\n")
    open(os.path.join(gen, "layer_type_enum.h"), "w").write("\n".join(enum) + "\n")


def extract_function(path, signature_prefix, nth=0):
    """Return the text of the nth top-level function whose first line starts with signature_prefix
    (up to and including the closing brace in column 0)."""
    lines = open(path).read().split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith(signature_prefix)]
    i = starts[nth]
    j = i
    while lines[j] != "}":
        j += 1
    return "\n".join(lines[i:j + 1]) + "\n", (i + 1, j + 1)


def gen_extracts(ref, gen):
    src = os.path.join(ref, "src")
    body, r1 = extract_function(os.path.join(src, "rife.cpp"), "int RIFE::process_cpu(")
    body2, r2 = extract_function(os.path.join(src, "rife.cpp"), "int RIFE::process_v4_cpu(")
    open(os.path.join(gen, "rife_cpu_extract.inc"), "w").write(
        "// extracted at build time from src/rife.cpp:%d-%d and :%d-%d -- do not commit\n" % (r1 + r2) + body + body2)
    wbody, r3 = extract_function(os.path.join(src, "warp.cpp"), "int Warp::forward(const std::vector<Mat>&")
    open(os.path.join(gen, "warp_cpu_extract.inc"), "w").write(
        "// extracted at build time from src/warp.cpp:%d-%d -- do not commit\n" % r3 + wbody)


def build_variant(ref, isa, jobs):
    ncnn_src = os.path.join(ref, "src", "ncnn", "src")
    gen = os.path.join(OUT, "gen_" + isa)
    obj = os.path.join(OUT, "obj_" + isa)
    os.makedirs(obj, exist_ok=True)
    gen_headers(ncnn_src, gen, isa)
    gen_extracts(ref, gen)
    srcs = [os.path.join(ncnn_src, f) for f in CORE]
    for cls in LAYERS:
        n = cls.lower()
        srcs.append(os.path.join(ncnn_src, "layer", n + ".cpp"))
        x = os.path.join(ncnn_src, "layer", "x86", n + "_x86.cpp")
        if os.path.exists(x):
            srcs.append(x)
    flags = ["-std=c++11", "-Ofast", "-ffast-math", "-fopenmp", "-fno-rtti", "-fno-exceptions", "-fPIC", "-w",
             "-I" + gen, "-I" + ncnn_src, "-I" + os.path.join(ncnn_src, "layer"),
             "-I" + os.path.join(ncnn_src, "layer", "x86")] + ISA_FLAGS[isa]

    def cc(s):
        o = os.path.join(obj, hashlib.md5(s.encode()).hexdigest()[:8] + "_" + os.path.basename(s) + ".o")
        if not (os.path.exists(o) and os.path.getmtime(o) > os.path.getmtime(s)):
            sh(["g++"] + flags + ["-c", s, "-o", o])
        return o

    with cf.ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(cc, srcs))
    main_o = os.path.join(obj, "ref_main.o")
    sh(["g++"] + flags + ["-c", os.path.join(HERE, "ref_main.cpp"), "-o", main_o])
    exe = os.path.join(OUT, "ref_rife_" + isa)
    sh(["g++", "-fopenmp", "-o", exe, main_o] + objs + ["-lpthread"])
    return exe


def copy_models(ref, names):
    for n in names:
        s = os.path.join(ref, "models", n)
        d = os.path.join(OUT, "models", n)
        if os.path.isdir(s) and not os.path.isdir(d):
            shutil.copytree(s, d)
    img = os.path.join(OUT, "images")
    if not os.path.isdir(img):
        shutil.copytree(os.path.join(ref, "images"), img)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--isa", default="avx2,avx512")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    # all eleven model directories the reference ships (src/main.cpp:658-683 sniffs them by name): 440 MB of weights that travel to
    # the GPU box with oracle/_ref, so `pytest -m gpu` there checks every one of them against the oracle instead of skipping seven
    ap.add_argument("--models", default="rife-v4.6,rife-v4,rife-v2.3,rife-anime,rife,rife-HD,rife-UHD,rife-v2,rife-v2.4,rife-v3.0,rife-v3.1")
    a = ap.parse_args()
    if not os.path.isdir(a.ref):
        print("build_ref: %s absent -- keeping prebuilt oracle/_ref as is" % a.ref)
        return 0
    os.makedirs(OUT, exist_ok=True)
    for isa in a.isa.split(","):
        exe = build_variant(a.ref, isa, a.jobs)
        print("built", exe)
    copy_models(a.ref, [m for m in a.models.split(",") if m])
    return 0


if __name__ == "__main__":
    sys.exit(main())
