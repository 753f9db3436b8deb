#!/usr/bin/env python3
"""Builds the reference's own CLI (src/main.cpp, UNMODIFIED) against this repo's `class RIFE` shim:
    rife-ncnn-vulkan_b200/host/_cli/rife-b200-cli
Proof of the drop-in boundary.  Needs /root/reference (main.cpp, stb headers, libwebp sources are
compiled where they lie); outputs are git-ignored.  libwebp is built from its C sources with gcc directly."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_cli")


def sh(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(" ".join(cmd[:6]) + " ...\n" + r.stdout[-3000:])
        raise SystemExit(1)


def main(ref="/root/reference"):
    src = os.path.join(ref, "src")
    if not os.path.exists(os.path.join(src, "main.cpp")):
        print("reference not mounted; nothing to do")
        return 0
    os.makedirs(OUT, exist_ok=True)
    webp = os.path.join(src, "libwebp")
    cs = []
    for sub in ("dec", "dsp", "enc", "utils"):
        cs += sorted(glob.glob(os.path.join(webp, "src", sub, "*.c")))
    lib = os.path.join(OUT, "libwebp_min.a")
    if not os.path.exists(lib):
        objs = []
        for c in cs:
            o = os.path.join(OUT, os.path.basename(os.path.dirname(c)) + "_" + os.path.basename(c)[:-2] + ".o")
            sh(["gcc", "-O2", "-w", "-fPIC", "-I" + webp, "-I" + os.path.join(webp, "src"), "-c", c, "-o", o])
            objs.append(o)
        sh(["ar", "rcs", lib] + objs)
    # main.cpp is copied next to nothing of ours: quoted includes resolve in its own directory first, so compile a copy
    tmp = os.path.join(OUT, "main.cpp")
    with open(os.path.join(src, "main.cpp")) as f, open(tmp, "w") as g:
        g.write(f.read())
    exe = os.path.join(OUT, "rife-b200-cli")
    sh(["g++", "-O2", "-std=c++11", "-fopenmp", "-w", "-I" + HERE, "-I" + os.path.join(HERE, "ncnn_compat"), "-I" + src, "-I" + os.path.join(webp, "src"),
        tmp, os.path.join(HERE, "rife.cpp"), lib, "-ldl", "-lpthread", "-o", exe])
    os.remove(tmp)  # the reference's source text does not stay in the repo tree
    print("built", exe)
    return 0


if __name__ == "__main__":
    sys.exit(main())
