"""The WHOLE engine on the host, against the oracle, without a GPU.  A test-only build links the engine's unmodified host code
(csrc/capi.cu, engine.cu, exec.cu, model.cpp) with host builds of its plain CUDA kernels (generic_kernels.cu, hbm_kernels.cu:
a block's threads are real threads, tests/emu/cuda_host_shim.h), a runtime whose device memory is host memory
(tests/emu/fake_cudart.cpp) and refusing stand-ins for the two tcgen05 pieces (tests/emu/emu_engine_stubs.cpp).  The package
is pointed at that library (RIFE_B200_LIB) in a child process and runs the same `parity.check_case` the GPU tests run, at
precision tier 0 (fp32 kernels everywhere): model loading, plan building and arena reuse, Split / Crop aliasing, the fused
conv epilogues, the v4 and the 3-net pipelines, the orientation fork / join of the TTA modes on helper lanes, staging of pageable
frames, the frame table -- everything but the tensor-core kernels -- is checked against the reference's own CPU path here.
Frames are tiny (a host 'block' is 256 OS threads meeting at barriers)."""
import json
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")
_LAUNCH = re.compile(r"(\w+(?:<[^<>;]*>)?)<<<(.*), (\d+), (\w+), st>>>\((.*)\);")


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    return _build(tmp_path_factory, False)


@pytest.fixture(scope="module")
def emu_lib_asan(tmp_path_factory):
    if not os.path.exists(_libasan()):
        pytest.skip("libasan not available")
    return _build(tmp_path_factory, True)


def _libasan():
    try:
        return subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
    except Exception:
        return ""


def _cpu_has(flag):
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return flag in line.split(":", 1)[1].split()
    except OSError:
        pass
    return False


def _rewrite(src):
    out, n = _LAUNCH.subn(lambda m: "emu_launch(dim3(%s), %s, %s, [&]() { %s(%s); });" % (m.group(2), m.group(3), m.group(4), m.group(1), m.group(5)), src)
    assert n == src.count("<<<"), (n, src.count("<<<"))
    return out


def _build(tmp_path_factory, asan):
    inc = "/usr/local/cuda/include"
    if os.environ.get("CUDA_HOME"):
        inc = os.path.join(os.environ["CUDA_HOME"], "include")
    if shutil.which("g++") is None or not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("g++ or CUDA headers not available")
    tsan = asan == "tsan"
    asan = asan is True
    d = str(tmp_path_factory.mktemp("emu_engine_tsan" if tsan else ("emu_engine_asan" if asan else "emu_engine")))
    gen = _rewrite(open(os.path.join(CSRC, "generic_kernels.cu")).read())
    assert gen.count("extern __shared__ float smem[];") == 1
    open(os.path.join(d, "generic_kernels_emu.inc"), "w").write(gen.replace("extern __shared__ float smem[];", "float* smem = reinterpret_cast<float*>(emu_dyn_smem);"))
    open(os.path.join(d, "hbm_kernels_emu.inc"), "w").write(_rewrite(open(os.path.join(CSRC, "hbm_kernels.cu")).read()))
    tc = open(os.path.join(CSRC, "tc_conv.cu")).read()
    a = tc.index("// ---- layout conversion kernels")
    helpers = [l for l in tc.splitlines() if l.startswith("__device__ __forceinline__ uint32_t pack2(") or l.startswith("int tc_conv_tile_rows(int N)")]
    assert len(helpers) == 2, helpers
    open(os.path.join(d, "tc_host_section.inc"), "w").write("\n".join(helpers) + "\n" + _rewrite(tc[a:]))
    open(os.path.join(d, "tu_generic.cpp"), "w").write('#include "cuda_host_shim.h"\n#include "generic_kernels_emu.inc"\n')
    open(os.path.join(d, "tu_hbm.cpp"), "w").write('#define EMU_ENGINE_BUILD 1\n#include "cuda_host_shim.h"\n#include "hbm_kernels_emu.inc"\n')
    # AddressSanitizer build: device memory is the (exact-size) host heap; fibers switch through ucontext there (ASan follows
    # swapcontext, not a hand-written switch)
    asan = ["-fsanitize=address", "-fno-omit-frame-pointer", "-g", "-DEMU_NO_FAST_SWITCH"] if asan else []
    if tsan:  # ThreadSanitizer over the engine's host threading: kernel launches do nothing (no fibers under TSan), native driver program
        asan = ["-fsanitize=thread", "-g", "-DEMU_SKIP_KERNELS"]
    flags = ["-O2", "-ffp-contract=off"] + asan + (["-mfma"] if _cpu_has("fma") else []) + ["-std=c++17", "-fPIC", "-pthread", "-w", "-I" + inc, "-I" + CSRC, "-I" + EMU, "-I" + d, "-I" + os.path.join(ROOT, "include")]
    units = [("tu_generic.cpp", os.path.join(d, "tu_generic.cpp")), ("tu_hbm.cpp", os.path.join(d, "tu_hbm.cpp")), ("stubs", os.path.join(EMU, "emu_engine_stubs.cpp")),
             ("fake_cudart", os.path.join(EMU, "fake_cudart.cpp"))] + [(f, os.path.join(CSRC, f)) for f in ("capi.cu", "engine.cu", "exec.cu", "model.cpp")]
    procs = []
    for name, path in units:
        obj = os.path.join(d, name.replace(".", "_") + ".o")
        extra = ["-Drife_b200_create=rife_b200_create_tier1"] if name == "capi.cu" else []  # see tests/emu/emu_engine_stubs.cpp
        procs.append((name, obj, subprocess.Popen(["g++"] + flags + extra + ["-x", "c++", "-c", path, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, obj, p in procs:
        out, _ = p.communicate()
        assert p.returncode == 0, (name, out[-4000:])
    if tsan:
        exe = os.path.join(d, "tsan_engine")
        r = subprocess.run(["g++"] + flags + ["-fsanitize=thread", "-o", exe, os.path.join(EMU, "tsan_engine_main.cpp")] + [o for _, o, _ in procs] + ["-ldl"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0 and "tsan" in r.stdout:
            pytest.skip("ThreadSanitizer runtime not available")
        assert r.returncode == 0, r.stdout[-6000:]
        return exe
    so = os.path.join(d, "librife_b200_hostemu.so")
    r = subprocess.run(["g++", "-shared", "-pthread"] + (["-fsanitize=address"] if asan else ["-Wl,--no-undefined"]) + ["-o", so] + [o for _, o, _ in procs] + ["-ldl"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-6000:]
    return so


_CHILD = r"""
import json, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import __graft_entry__ as g
import parity
pkg = g.load_package()
assert pkg.LIB_PATH == %(so)r, pkg.LIB_PATH
out = []
for case in json.loads(sys.argv[1]):
    model, w, h = case.pop("model"), case.pop("w"), case.pop("h")
    opts = {"precision": 0}
    opts.update(case.pop("options", {}))
    res = parity.check_case(pkg, model, w, h, options=opts, **case)
    out.append(res)
print("RESULT " + json.dumps(out))
"""


def _run(so, cases, timeout=1500, asan=False):
    env = dict(os.environ, RIFE_B200_LIB=so)
    if asan:  # the interpreter is not an ASan build: the runtime has to be loaded first
        env.update(LD_PRELOAD=_libasan(), ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    code = _CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests"), "so": so}
    r = subprocess.run([sys.executable, "-c", code, json.dumps(cases)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert r.returncode == 0 and "ERROR: AddressSanitizer" not in r.stdout, r.stdout[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, r.stdout[-2000:]
    return json.loads(line[0][7:])


def _ok(res):
    assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50 and res["out_std"] > 5, res


def test_v4_family_plain(emu_lib):
    """rife-v4.6 (also a ragged size: the padded result is cropped) and rife-v4 with an off-centre timestep"""
    for res in _run(emu_lib, [{"model": "rife-v4.6", "w": 64, "h": 64}, {"model": "rife-v4.6", "w": 40, "h": 36, "timestep": 0.3},
                              {"model": "rife-v4", "w": 64, "h": 32, "timestep": 0.25}]):
        _ok(res)


def test_v46_tta_modes(emu_lib):
    """-z, and -x -z: eight orientations dealt to helper lanes between the flow-averaging points, 16 inputs to the TTA postproc"""
    for res in _run(emu_lib, [{"model": "rife-v4.6", "w": 48, "h": 32, "tta_temporal": True}, {"model": "rife-v4.6", "w": 32, "h": 32, "tta": True, "tta_temporal": True}]):
        _ok(res)


ALL_MODELS = ["rife", "rife-HD", "rife-UHD", "rife-anime", "rife-v2", "rife-v2.3", "rife-v2.4", "rife-v3.0", "rife-v3.1", "rife-v4", "rife-v4.6"]


def test_model_families_on_the_host_build(emu_lib):
    """The model directories the reference ships (src/main.cpp:658-683), one small frame each: flownet + contextnet + fusionnet
    families with 5x5 convolutions, SE blocks (pooling, inner product, broadcast multiply), PReLU.  By default one directory per
    graph family (the v4 layouts and rife-v2.3 run in the tests above); RIFE_EMU_FULL=1: all eleven (a run of all eleven: 9 bit-identical,
    rife-v2 and rife-v3.0 with single 1-LSB flips, 78 / 83 dB)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    pick = ALL_MODELS if os.environ.get("RIFE_EMU_FULL") else ["rife-HD", "rife-anime", "rife-v3.1"]
    have = [m for m in pick if parity.model_dir(m)]
    assert have
    cases = [dict({"model": m, "w": 32, "h": 32}, **({"timestep": 0.4} if m.startswith("rife-v4") else {})) for m in have]
    for m, res in zip(have, _run(emu_lib, cases, timeout=3000)):
        assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50 and res["out_std"] > 5, (m, res)


def test_uhd_mode_on_the_host_build(emu_lib):
    """-u: the flownet runs on the half-size frame, the flow is scaled back up (rife.cpp:2212-2229)"""
    for res in _run(emu_lib, [{"model": "rife-v2.3", "w": 64, "h": 64, "uhd": True}]):
        _ok(res)


@pytest.mark.skipif(not os.environ.get("RIFE_EMU_FULL"), reason="several minutes on the host; RIFE_EMU_FULL=1 runs it (results of such a run: profiles/README.md)")
def test_spatial_tta_of_the_three_net_families(emu_lib):
    for res in _run(emu_lib, [{"model": "rife", "w": 32, "h": 32, "tta": True, "tta_temporal": True}, {"model": "rife-v2.3", "w": 32, "h": 32, "tta": True},
                              {"model": "rife-v2.3", "w": 32, "h": 32, "tta": True, "tta_temporal": True}], timeout=3000):
        _ok(res)


_API_CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import __graft_entry__ as g
import parity
pkg = g.load_package()
assert pkg.LIB_PATH == %(so)r
w, h = 32, 32
frames = parity.synth.stream(0, 4, w, h)
r = pkg.RIFE(0, False, False, False, 1, False, True)
r.load(parity.model_dir("rife-v4.6"))
r.set_option("precision", 0)
singles = [r.process(frames[i], frames[i + 1], 0.5) for i in range(3)]
h2d0 = pkg.copy_bytes()[0]
outs = [np.empty_like(frames[0]) for _ in range(3)]
r.process_batch_ptr([f.ctypes.data for f in frames[:3]], [f.ctypes.data for f in frames[1:]], w, h, [0.5, 1.0, 0.5], [o.ctypes.data for o in outs])
h2d = pkg.copy_bytes()[0] - h2d0
ok_batch = bool(np.array_equal(outs[0], singles[0]) and np.array_equal(outs[1], frames[2]) and np.array_equal(outs[2], singles[2]))
r.set_option("frame_cache", 1)
again = [r.process(frames[i], frames[i + 1], 0.5) for i in range(3)]
hits = r.get_option("frame_cache_hits")
err = None
try:
    r.process_batch_ptr([frames[0].ctypes.data, 0], [frames[1].ctypes.data, frames[2].ctypes.data], w, h, [0.5, 0.5], [outs[0].ctypes.data, outs[1].ctypes.data])
except pkg.RifeError as e:
    err = str(e)
after = r.process(frames[0], frames[1], 0.5)
r.close()
print("RESULT " + json.dumps({"ok_batch": ok_batch, "h2d_frames": h2d / (w * h * 3), "cache_equal": all(bool(np.array_equal(a, b)) for a, b in zip(singles, again)),
                               "hits": hits, "null_frame_error": err, "usable_after_error": bool(np.array_equal(after, singles[0]))}))
"""


def test_batch_call_frame_table_and_error_path(emu_lib):
    """process_batch on host buffers (every frame of the call uploaded once, a timestep edge inside the batch), the cross-call frame
    cache, and a refused call (null frame) that leaves the handle usable -- the engine's host logic, on the host"""
    env = dict(os.environ, RIFE_B200_LIB=emu_lib)
    code = _API_CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests"), "so": emu_lib}
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    assert res["ok_batch"] and res["cache_equal"] and res["usable_after_error"], res
    assert res["h2d_frames"] == 4 and res["hits"] == 2, res  # pairs (0,1) and (2,3) of the batch: four uploads (the t = 1 pair is a host copy); the cache finds frames 1 and 2 again
    assert res["null_frame_error"], res


def test_address_sanitizer_build(emu_lib_asan):
    """The same host build under AddressSanitizer.  Device allocations are exact-size heap blocks there, so an access of an
    emulated kernel or of the engine's host code past an ALLOCATION -- a frame, a per-lane buffer, a weight blob, an executor arena
    as a whole (tensors inside one arena are not separated) -- is a report (checked by injection: a postproc that writes its last
    row 4 bytes late aborts with heap-buffer-overflow).  A ragged frame through the v4 pipeline -- partial tiles in every kernel --
    must finish without a report and still match the oracle.  RIFE_EMU_FULL=1 adds
    the TTA, UHD and 3-net cases (10+ minutes; a clean run of all of them is recorded in profiles/README.md)."""
    cases = [{"model": "rife-v4.6", "w": 40, "h": 36, "timestep": 0.3}]
    if os.environ.get("RIFE_EMU_FULL"):
        cases += [{"model": "rife-v4.6", "w": 32, "h": 32, "tta": True, "tta_temporal": True}, {"model": "rife-v2.3", "w": 64, "h": 64, "uhd": True},
                  {"model": "rife-anime", "w": 32, "h": 32}, {"model": "rife-v2.3", "w": 32, "h": 32, "tta": True, "tta_temporal": True}]
    for res in _run(emu_lib_asan, cases, timeout=3600, asan=True):
        _ok(res)


def test_reference_cli_on_the_host_build(emu_lib, tmp_path):
    """The reference's UNMODIFIED src/main.cpp linked against the `class RIFE` shim (host/_cli/rife-b200-cli, built by build() when
    /root/reference is present), in directory mode with its load / proc / save threads (`-j 1:2:2`: two threads call RIFE::process
    on one handle), against the host build: the drop-in boundary end to end -- dlopen of the library, the ncnn-namespace shim, the
    request combiner, PNG in and out -- without a GPU.  Every written frame against the oracle (src/main.cpp:712-731 schedule)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    cli = os.path.join(ROOT, "rife-ncnn-vulkan_b200", "host", "_cli", "rife-b200-cli")
    md = parity.model_dir("rife-v4.6")
    if not os.path.exists(cli) or md is None or not os.path.exists(os.path.join(md, "flownet.bin")):
        pytest.skip("host/_cli/rife-b200-cli (needs /root/reference at build time) or the reference model files are missing")
    try:
        from PIL import Image
    except Exception:
        pytest.skip("PIL not available")
    from test_cli_dropin_gpu import _dir_schedule
    w, h, count = 32, 24, 3  # w % 32 == 0: for other widths the reference CPU binary shears the frame (output crop)
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir()
    outd.mkdir()
    frames = parity.synth.stream(0, count, w, h)
    for k, f in enumerate(frames):
        Image.fromarray(f).save(str(ind / ("%08d.png" % (k + 1))))
    env = dict(os.environ, RIFE_B200_LIB=emu_lib)
    r = subprocess.run([cli, "-i", str(ind), "-o", str(outd), "-m", md, "-g", "0", "-j", "1:2:2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    for i, (sx, t) in enumerate(_dir_schedule(count, 2 * count)):
        p = outd / ("%08d.png" % (i + 1))
        assert p.exists(), (i, r.stderr[-1000:])
        got = np.array(Image.open(str(p)).convert("RGB"))
        if t == 0.0:
            assert np.array_equal(got, frames[sx]), i
        elif t == 1.0:
            assert np.array_equal(got, frames[sx + 1]), i
        else:
            ref, _ = parity.run_oracle("rife-v4.6", frames[sx], frames[sx + 1], t)
            res = parity.compare(got, ref)
            assert res["max_abs_diff"] <= 1 and res["psnr_db"] > 50, (i, sx, t, res)


_FUZZ_CHILD = r"""
import json, os, random, shutil, sys, tempfile
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import __graft_entry__ as g
import parity
pkg = g.load_package()
assert pkg.LIB_PATH == %(so)r
src = parity.model_dir("rife-v4.6")
lines = open(os.path.join(src, "flownet.param")).read().splitlines()
a, b = parity.synth.pair(32, 32)
TYPES = ["Convolution", "Deconvolution", "BinaryOp", "ReLU", "Concat", "Split", "Crop", "Interp", "PixelShuffle", "Sigmoid", "Eltwise", "rife.Warp", "Input", "PReLU",
         "InnerProduct", "Pooling"]


def mutate(seed):
    rnd = random.Random(seed)
    L = list(lines)
    i = rnd.randrange(2, len(L))
    tok = L[i].split()
    kind = rnd.randrange(6)
    if kind == 0:    # a parameter value
        idx = [k for k, t in enumerate(tok) if "=" in t and k >= 4]
        if idx:
            k = rnd.choice(idx)
            tok[k] = tok[k].split("=", 1)[0] + "=" + rnd.choice(["0", "-1", "1", "2", "3", "5", "7", "16", "64", "100", "512", "-23310", "1.5"])
    elif kind == 1:  # a blob reference
        n = int(tok[2]) + int(tok[3])
        if n:
            tok[4 + rnd.randrange(n)] = str(rnd.randrange(0, 276))
    elif kind == 2:  # the layer type
        tok[0] = rnd.choice(TYPES)
    elif kind == 3:
        tok[2] = str(max(0, int(tok[2]) + rnd.choice([-1, 1])))
    elif kind == 4:
        L.pop(i)
        tok = None
    else:
        tok[3] = str(max(0, int(tok[3]) + rnd.choice([-1, 1])))
    if tok is not None:
        L[i] = " ".join(tok)
    return L


out = {}
for seed in json.loads(sys.argv[1]):
    d = tempfile.mkdtemp()
    open(os.path.join(d, "flownet.param"), "w").write("\n".join(mutate(seed)) + "\n")
    os.symlink(os.path.join(src, "flownet.bin"), os.path.join(d, "flownet.bin"))
    print("SEED %%d" %% seed, flush=True)
    r = pkg.RIFE(0, False, False, False, 1, False, True)
    try:
        r.load(d)
        try:
            r.process(a, b, 0.5)
            out[seed] = "ok"
        except pkg.RifeError as e:
            out[seed] = "process refused"
    except pkg.RifeError as e:
        out[seed] = "load refused"
    r.close()
    shutil.rmtree(d)
print("RESULT " + json.dumps(out))
"""


def test_damaged_graphs_are_refused_not_executed(emu_lib_asan):
    """One random edit per case to rife-v4.6's flownet.param (a parameter value, a blob reference, a layer type, an input / output
    count, a dropped line), weights untouched; then load + one frame on the AddressSanitizer host build.  Every case must end in a
    result or in an error return -- never in a report or a crash.  (This found two reads the plan builder did not guard: a layer
    with fewer inputs than its type takes -- csrc/exec.cu indexed bottoms by position -- and a PReLU without slope data, whose
    kernel would have dereferenced a null pointer on the device.)  RIFE_EMU_FULL=1: 400 cases instead of 24."""
    seeds = list(range(400 if os.environ.get("RIFE_EMU_FULL") else 24))
    env = dict(os.environ, RIFE_B200_LIB=emu_lib_asan, LD_PRELOAD=_libasan(), ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:allocator_may_return_null=1")
    code = _FUZZ_CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests"), "so": emu_lib_asan}
    r = subprocess.run([sys.executable, "-c", code, json.dumps(seeds)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3600)
    last = [l for l in r.stdout.splitlines() if l.startswith("SEED ")][-1:]
    assert r.returncode == 0 and "ERROR: AddressSanitizer" not in r.stdout, (last, r.stdout[-3000:])
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    assert len(res) == len(seeds)
    kinds = set(res.values())
    assert kinds <= {"ok", "process refused", "load refused"} and "process refused" in kinds and "load refused" in kinds, kinds


def test_gpu_suite_cases_that_need_no_tensor_cores(emu_lib):
    """Cases of tests/test_parity_gpu.py that do not depend on the tcgen05 path, run UNCHANGED against the host build (the package
    takes its library from RIFE_B200_LIB): the t = 0 / 1 copies, the ragged-width crop both ways (the padded-crop default against
    the restatement and against the zero-padded-by-the-caller property; option cpu_crop_quirk against the reference binary), a null
    frame inside a batch.  RIFE_EMU_FULL=1 adds the 100x70 and rife-v2.3 variants."""
    k = "timestep_edges or null_frame or (ragged_widths and 90-50 and not v2.3)"
    if os.environ.get("RIFE_EMU_FULL"):
        k = "timestep_edges or null_frame or ragged_widths"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_parity_gpu.py"), "-q", "-m", "gpu", "-k", k, "-p", "no:cacheprovider"],
                       env=dict(os.environ, RIFE_B200_LIB=emu_lib), cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3000)
    tail = r.stdout.strip().splitlines()[-1]
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail and "skipped" not in tail, r.stdout[-3000:]
    assert int(tail.split(" passed")[0].split()[-1]) >= 5, tail


_ABI_CHILD = r"""
ROOT_DIR, TESTS_DIR = %(root)r, %(tests)r
import sys, os, ctypes, json
import numpy as np
sys.path.insert(0, ROOT_DIR); sys.path.insert(0, TESTS_DIR)
import __graft_entry__ as g
import parity
pkg = g.load_package()
L = pkg.lib()
vp = ctypes.c_void_p
res = {}
def call(name, fn):
    print("CALL", name, flush=True)
    try:
        res[name] = fn()
    except Exception as e:
        res[name] = "py:" + type(e).__name__
h = vp()
res["create_cpu"] = L.rife_b200_create(ctypes.byref(h), -1, 0, 0, 0, 1, 0, 1)
res["create_bad_gpu"] = L.rife_b200_create(ctypes.byref(h), 7, 0, 0, 0, 1, 0, 1)
res["create_null"] = L.rife_b200_create(None, 0, 0, 0, 0, 1, 0, 1)
assert L.rife_b200_create(ctypes.byref(h), 0, 0, 0, 0, 1, 0, 1) == 0
a, b = parity.synth.pair(32, 32)
o = np.empty_like(a)
P = lambda x: x.ctypes.data
call("process_before_load", lambda: L.rife_b200_process(h, P(a), P(b), 32, 32, 0.5, P(o)))
call("load_null", lambda: L.rife_b200_load(h, None))
call("load_missing", lambda: L.rife_b200_load(h, b"/nonexistent/dir"))
call("load_null_handle", lambda: L.rife_b200_load(None, parity.model_dir("rife-v4.6").encode()))
assert L.rife_b200_load(h, parity.model_dir("rife-v4.6").encode()) == 0
for name, args in {"w0": (0, 32), "h0": (32, 0), "wneg": (-5, 32), "hneg": (32, -1)}.items():
    call("process_" + name, lambda args=args: L.rife_b200_process(h, P(a), P(b), args[0], args[1], 0.5, P(o)))
call("process_null_in0", lambda: L.rife_b200_process(h, None, P(b), 32, 32, 0.5, P(o)))
call("process_null_out", lambda: L.rife_b200_process(h, P(a), P(b), 32, 32, 0.5, None))
call("process_null_handle", lambda: L.rife_b200_process(None, P(a), P(b), 32, 32, 0.5, P(o)))
call("process_nan_t", lambda: L.rife_b200_process(h, P(a), P(b), 32, 32, float("nan"), P(o)))
call("process_t_out_of_range", lambda: L.rife_b200_process(h, P(a), P(b), 32, 32, 7.5, P(o)))
arr = (vp * 2)(P(a), P(a)); arrb = (vp * 2)(P(b), P(b)); arro = (vp * 2)(P(o), P(o)); ts = (ctypes.c_float * 2)(0.5, 0.5)
call("batch_n_negative", lambda: L.rife_b200_process_batch(h, -3, arr, arrb, 32, 32, ts, arro))
call("batch_n_zero", lambda: L.rife_b200_process_batch(h, 0, arr, arrb, 32, 32, ts, arro))
call("batch_null_arrays", lambda: L.rife_b200_process_batch(h, 2, None, arrb, 32, 32, ts, arro))
call("batch_null_ts", lambda: L.rife_b200_process_batch(h, 2, arr, arrb, 32, 32, None, arro))
v = ctypes.c_int()
call("get_unknown", lambda: L.rife_b200_get_option(h, b"no_such_option", ctypes.byref(v)))
call("get_null_key", lambda: L.rife_b200_get_option(h, None, ctypes.byref(v)))
call("get_null_out", lambda: L.rife_b200_get_option(h, b"lanes", None))
call("set_unknown", lambda: L.rife_b200_set_option(h, b"no_such_option", 1))
for key, val in (("lanes", 0), ("lanes", -4), ("lanes", 1000), ("batch", -1), ("batch", 999), ("precision", 9), ("precision", -2), ("plain_blocks", -1), ("recompute_fm", 77)):
    call("set_%%s_%%d" %% (key, val), lambda key=key, val=val: L.rife_b200_set_option(h, key.encode(), val))
call("process_after_option_abuse", lambda: L.rife_b200_process(h, P(a), P(b), 32, 32, 0.5, P(o)))
sz = ctypes.c_size_t()
call("weights_size_null", lambda: L.rife_b200_weights_size(h, None))
call("weights_export_small", lambda: (L.rife_b200_weights_size(h, ctypes.byref(sz)), L.rife_b200_weights_export(h, P(o), 16))[1])
call("load_packed_null", lambda: L.rife_b200_load_packed(h, None, 100))
call("load_packed_zero", lambda: L.rife_b200_load_packed(h, P(o), 0))
buf = ctypes.create_string_buffer(8)
call("stage_report_tiny", lambda: L.rife_b200_stage_report(h, buf, 8))
call("stage_report_null", lambda: L.rife_b200_stage_report(h, None, 0))
call("last_error_null", lambda: bool(L.rife_b200_last_error(None) is not None))
call("forget_null", lambda: L.rife_b200_forget_frames(None))
call("debug_hbm_bad", lambda: L.rife_b200_debug_hbm(0, None, 9, 32, 32, 3, 0, P(a), None, P(o)))
call("debug_pack_bad", lambda: L.rife_b200_debug_pack_weights(5, 8, 16, 16, 0, 0, P(o), P(o), 3))
ok = L.rife_b200_process(h, P(a), P(b), 32, 32, 0.5, P(o))
res["final_process"] = ok
L.rife_b200_destroy(h)
L.rife_b200_destroy(None)
print("RESULT " + json.dumps(res))
"""


def test_c_abi_refuses_bad_arguments(emu_lib_asan):
    """Every entry point of include/rife_b200.h with the arguments a careless binding could pass -- null handles and pointers,
    zero / negative sizes, a missing model directory, process before load, unknown and out-of-range options, undersized export
    buffers, a damaged packed model -- on the AddressSanitizer host build: an error code every time, no report, and the handle
    interpolates a frame afterwards."""
    env = dict(os.environ, RIFE_B200_LIB=emu_lib_asan, LD_PRELOAD=_libasan(), ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:allocator_may_return_null=1")
    code = _ABI_CHILD % {"root": ROOT, "tests": os.path.join(ROOT, "tests")}
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)
    last = [l for l in r.stdout.splitlines() if l.startswith("CALL ")][-1:]
    assert r.returncode == 0 and "ERROR: AddressSanitizer" not in r.stdout, (last, r.stdout[-3000:])
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    refused = ["create_cpu", "create_bad_gpu", "create_null", "process_before_load", "load_null", "load_missing", "load_null_handle", "process_w0", "process_h0", "process_wneg",
               "process_hneg", "process_null_in0", "process_null_out", "process_null_handle", "batch_n_negative", "batch_null_arrays", "batch_null_ts", "get_unknown",
               "get_null_key", "get_null_out", "set_unknown", "weights_size_null", "weights_export_small", "load_packed_null", "load_packed_zero", "stage_report_null",
               "forget_null", "debug_hbm_bad", "debug_pack_bad"]
    for k in refused:
        assert isinstance(res[k], int) and res[k] < 0, (k, res[k])
    assert res["create_cpu"] == -1 and res["process_before_load"] == -4 and res["load_missing"] == -3  # the codes include/rife_b200.h documents
    assert res["batch_n_zero"] == 0 and res["process_after_option_abuse"] == 0 and res["final_process"] == 0 and res["last_error_null"] is True, res


def test_engine_threading_under_thread_sanitizer(tmp_path_factory):
    """tests/emu/tsan_engine_main.cpp: six threads call rife_b200_process on one handle with pageable frames (staging slots, the
    request combiner, the lock-free option snapshot, per-thread error text) while a seventh flips options and drops the frame cache,
    on a ThreadSanitizer build of the engine's host code whose kernel launches do nothing.  No report, no failed call."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    exe = _build(tmp_path_factory, "tsan")
    r = subprocess.run([exe, parity.model_dir("rife-v4.6")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    if r.returncode != 0 and "ThreadSanitizer" not in r.stdout and "TSAN-ENGINE" not in r.stdout:
        pytest.skip("ThreadSanitizer cannot run in this sandbox: " + r.stdout[-200:])
    assert r.returncode == 0 and "TSAN-ENGINE failures=0" in r.stdout and "WARNING: ThreadSanitizer" not in r.stdout, r.stdout[-4000:]
