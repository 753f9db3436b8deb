#!/usr/bin/env python3
"""bench.py -- interpolated frames/sec of the RIFE hot path (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p|4k] [--impl ours|reference]

A step = one pass of the hot path over one batch of PAIRS_PER_STEP synthetic frame pairs (rife-v4.6, t = 0.5).
  value     whole-job frames/s with the frames resident in HBM (rife_b200_process_device), CUDA-event timed on the
            stream the kernels are launched on, max over ranks.
  e2e       the same metric through the reference-facing call with HOST buffers (rife_b200_process_batch, pinned
            memory): H2D of two u8 frames + D2H of one per pair inside the timed region.
  roofline  the dominant kernel (tcgen05 conv3x3 64->64 at quarter resolution, 44 % of the model's FLOPs) timed alone
            with CUDA events; achieved = 2*9*Cin*Cout*H*W FLOP per launch / time; peak = measured cuBLAS bf16 TF/s.
  cpu_baseline  the reference's own CPU path (oracle/_ref: its rife.cpp CPU functions + vendored ncnn) on the host
            cores, bounded sample, rank 0 only.
--impl reference times that CPU path alone on the same workload (the driver computes the ratio).
Multi-GPU: one process per GPU (torchrun); frame pairs are independent, so ranks share nothing after rank 0
broadcasts the packed model over NCCL; scaling is weak (fixed pairs per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {"1080p": (1920, 1080, "rife-v4.6 1920x1080 synthetic frame-pair stream (BASELINE configs[1])"),
             "4k": (3840, 2160, "rife-v4.6 3840x2160 UHD-flag stream (BASELINE configs[2]; -u is a no-op for v4 nets)")}
GFLOP_PER_FRAME = {"1080p": 175.2, "4k": 701.0}
PAIRS_PER_STEP = int(os.environ.get("RIFE_BENCH_PAIRS", "0"))  # 0 = default for the workload (128 pairs of 1080p, 32 of 4K); profiling runs shrink the step
DISTINCT_FRAMES = 9  # consecutive frames of the synthetic stream; pairs cycle through them
MODEL = "rife-v4.6"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML in-process, every 5 ms;
    falls back to nvidia-smi when pynvml is unavailable)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, idx):
        super().__init__(daemon=True)
        self.idx, self.sm, self.mx, self.reasons, self.stop_flag = idx, [], None, set(), False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(idx))
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(idx):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            parts = [p for p in vis.split(",") if p.strip()]
            if idx < len(parts) and parts[idx].strip().isdigit():
                return int(parts[idx])
        return idx

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml:
                    self.sm.append(self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
                    try:
                        r = self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                    time.sleep(0.005)
                else:
                    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                    o = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                    if o:
                        r = [x.strip() for x in o.split(",")]
                        if r[0].isdigit():
                            self.sm.append(int(r[0]))
                        if r[1].isdigit():
                            self.mx = int(r[1])
                        for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                            if v.lower().startswith("active"):
                                self.reasons.add(name)
                    time.sleep(0.1)
            except Exception:
                time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": len(sm)}


def cpu_reference_fps(workload, frames, threads=None, warmup=1):
    """Times the reference's CPU path on `frames` frames of the workload; returns dict for the JSON line."""
    import parity
    w, h, _ = WORKLOADS[workload]
    a, b = parity.synth.pair(w, h)
    ncpu = os.cpu_count() or 1
    if threads is None:
        # the reference's OpenMP scaling is not monotonic (measured on the 128-thread B200 host: 16 threads 1.11 s/frame,
        # 32: 1.27, 64: 2.2, 128: 10.5 at 1080p -- profiles/r1_cpu_thread_sweep.txt): pick the best of a short sweep
        best = None
        for t in sorted({min(ncpu, c) for c in (8, 16, 32)}):
            _, i = parity.run_oracle(MODEL, a, b, 0.5, threads=t, repeat=1, warmup=0)
            if best is None or i["sec_per_frame"][0] < best[1]:
                best = (t, i["sec_per_frame"][0])
        threads = best[0]
    _, info = parity.run_oracle(MODEL, a, b, 0.5, threads=threads, repeat=frames, warmup=warmup)
    secs = info["sec_per_frame"]
    fps = len(secs) / sum(secs)
    return {"value": fps, "unit": "frames/s", "cores": threads, "kind": info["kind"],
            "sample": "%d frames of %s after %d warm-up (%.2f s/frame)" % (len(secs), workload, warmup, sum(secs) / len(secs))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w, h, desc = WORKLOADS[args.workload]
    frames = 2 if args.workload == "1080p" else 1
    t0 = time.time()
    per_step = []
    base = None
    first = cpu_reference_fps(args.workload, 1, warmup=0)  # also picks the thread count
    for _ in range(args.steps):
        base = cpu_reference_fps(args.workload, frames, threads=first["cores"], warmup=0)
        per_step.append(base["value"])
        if time.time() - t0 > 240:
            break
    fps = sum(per_step) / len(per_step)
    base["value"] = fps
    base["sample"] = "each step = %d frames of %s; %d steps" % (frames, args.workload, len(per_step))
    line = {"impl": "reference", "metric": "interpolated frames/sec (rife-v4.6)", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": len(per_step), "warmup": args.warmup, "ms_per_step": 1000.0 * frames / fps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "timestep": 0.5},
            "cpu_baseline": base, "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="1080p", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", type=int, default=1)
    ap.add_argument("--lanes", type=int, default=2)
    ap.add_argument("--plain-blocks", type=int, default=-1, help="bit mask of IFBlocks whose residual chain uses plain fp16 activations (-1 = library default)")
    ap.add_argument("--recompute-fm", type=int, default=-1, help="fused path: rebuild the full-resolution flow / mask planes instead of storing them (0, 1, 2; -1 = library default)")
    ap.add_argument("--batch", type=int, default=0, help="pairs per lock-step batch on the fused path (0 = auto from the frame size)")
    ap.add_argument("--model", default=MODEL, help="model directory name (default rife-v4.6 = the BASELINE metric; others are side measurements)")
    ap.add_argument("--tta", action="store_true")
    ap.add_argument("--tta-temporal", action="store_true")
    ap.add_argument("--timestep", type=float, default=0.5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global PAIRS_PER_STEP
    if PAIRS_PER_STEP <= 0:
        PAIRS_PER_STEP = 32 if args.workload == "4k" else 128
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import __graft_entry__ as g
    import parity
    pkg = g.load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w, h, desc = WORKLOADS[args.workload]
    md = parity.model_dir(args.model)
    if md is None:
        raise SystemExit("model %s not found (oracle/_ref/models or tests/models)" % args.model)
    v2, v4 = pkg.family_flags(args.model)
    eng = pkg.RIFE(local, args.tta, args.tta_temporal, args.workload == "4k" and v4, 1, v2, v4)
    # weights: rank 0 reads the model directory, everyone else receives the packed blob over NCCL (NVLink)
    if world > 1:
        sys.path.insert(0, os.path.join(ROOT, "rife-ncnn-vulkan_b200"))
        import dist_util
        blob = None
        if rank == 0:
            eng.load(md)
            blob = eng.export_weights()
        blob = dist_util.broadcast_blob(blob, dist, device=torch.device("cuda", local))
        if rank != 0:
            eng.load_packed(blob)
    else:
        eng.load(md)
    eng.set_option("precision", args.precision)
    eng.set_option("lanes", args.lanes)
    eng.set_option("batch", args.batch)
    if args.plain_blocks >= 0:
        eng.set_option("plain_blocks", args.plain_blocks)
    if args.recompute_fm >= 0:
        eng.set_option("recompute_fm", args.recompute_fm)

    # synthetic frames: a short stream, distinct per rank; PAIRS_PER_STEP consecutive pairs per step
    nframes = min(PAIRS_PER_STEP, DISTINCT_FRAMES - 1) + 1
    frames = [parity.synth.frame(k, w, h, seed=rank) for k in range(nframes)]
    host = [torch.from_numpy(f).pin_memory() for f in frames]
    dev = [t.cuda(non_blocking=True) for t in host]
    out_dev = [torch.empty_like(dev[0]) for _ in range(PAIRS_PER_STEP)]
    out_host = [torch.empty_like(host[0]).pin_memory() for _ in range(PAIRS_PER_STEP)]
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_option("async", 1)
    torch.cuda.synchronize()

    d_in0 = [dev[i % (nframes - 1)].data_ptr() for i in range(PAIRS_PER_STEP)]
    d_in1 = [dev[i % (nframes - 1) + 1].data_ptr() for i in range(PAIRS_PER_STEP)]
    d_out = [t.data_ptr() for t in out_dev]

    def step_device():
        eng.process_batch_ptr(d_in0, d_in1, w, h, [args.timestep] * PAIRS_PER_STEP, d_out, device=True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        barrier()
        launches0 = pkg.launch_count()
        sampler = ClockSampler(local)
        sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            l2_flush.zero_()  # flush L2 between timed iterations (outside the event bracket)
            ev[k][0].record(stream)
            step_device()
            ev[k][1].record(stream)
        barrier()
        launches = pkg.launch_count() - launches0
    ms_dev = sum(a.elapsed_time(b) for a, b in ev)
    t_local = torch.tensor([ms_dev], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    ms_total = float(t_local.item())
    value = world * PAIRS_PER_STEP * args.steps / (ms_total / 1000.0)

    # e2e: host buffers through the batch call (H2D + compute + D2H pipelined inside the library)
    eng.set_stream(0)
    eng.set_option("async", 0)
    in0 = [host[i % (nframes - 1)].data_ptr() for i in range(PAIRS_PER_STEP)]
    in1 = [host[i % (nframes - 1) + 1].data_ptr() for i in range(PAIRS_PER_STEP)]
    outp = [t.data_ptr() for t in out_host]
    ts = [args.timestep] * PAIRS_PER_STEP
    for _ in range(2):
        eng.process_batch_ptr(in0, in1, w, h, ts, outp)
    barrier()
    cb0 = pkg.copy_bytes()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.process_batch_ptr(in0, in1, w, h, ts, outp)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    cb1 = pkg.copy_bytes()
    h2d_step, d2h_step = (cb1[0] - cb0[0]) // args.steps, (cb1[1] - cb0[1]) // args.steps
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = world * PAIRS_PER_STEP * args.steps / float(t_e2e.item())
    clocks = sampler.summary()
    checksum = int(out_host[0].to(torch.int64).sum().item())

    # roofline of the dominant kernel: conv3x3 64->64 (+res +leaky) at (h/4) x (w/4), split-fp16 operands count double
    burst, sustained, hbm, how = measured_peaks()
    hp, wp = (h + 31) // 32 * 32, (w + 31) // 32 * 32
    ch, cw = hp // 4, wp // 4
    pmask = eng.get_option("plain_blocks") if eng.get_option("fast_active") else 0
    split = 1 if (args.precision == 1 and not (pmask & 8)) else 0  # block 3's residual chain: plain fp16 when bit 3 is set
    iters = 20
    # images per launch exactly as in the timed step (the lock-step batch of the fused path; 1 on the generic path)
    kb = 1
    if eng.get_option("fast_active"):
        kb = args.batch if args.batch > 0 else max(1, min(8, (2 * 3840 * 2176) // (hp * wp)))
        kb = min(kb, PAIRS_PER_STEP)
    with torch.cuda.stream(stream):
        pkg.bench_conv(stream.cuda_stream, 64, 64, ch, cw, split, 3, gpuid=local, batch=kb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pkg.bench_conv(stream.cuda_stream, 64, 64, ch, cw, split, iters, gpuid=local, batch=kb)
        e1.record(stream)
        torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / iters
    flop = 2.0 * 9 * 64 * 64 * ch * cw * kb  # algorithmic FLOPs per launch; a hi+lo split issues 2x this on the tensor pipe
    achieved = flop / (k_ms * 1e-3) / 1e12
    # DRAM bytes per launch from the committed `ncu --set full` capture of this kernel (profiles/README.md)
    # (single-image captures, scaled by the images per launch: the kernel re-reads nothing across images)
    # keyed "<w>x<h>x<images>_<split|plain>" -> dram__bytes_read.sum + dram__bytes_write.sum of one launch
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "conv64_dram_traffic.json")))
        traffic = tj.get("%dx%dx%d_%s" % (cw, ch, kb, "split" if split else "plain"))
        if traffic is None:  # single-image capture scaled by the images per launch (the kernel re-reads nothing across images)
            one = tj.get("%dx%dx1_%s" % (cw, ch, "split" if split else "plain"))
            traffic = one * kb if one else None
    except (OSError, ValueError):
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst, "traffic": traffic,
                "kernel": "tc_conv3x3_kernel<64,4,3,9> %d x %dx%d" % (kb, cw, ch), "images_per_launch": kb, "us_per_launch": k_ms * 1000.0, "peak_source": how + " bf16 burst",
                "tensor_issue_multiplier": 2 if split else 1}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and args.model == MODEL and not (args.tta or args.tta_temporal):
            try:
                cpu = cpu_reference_fps(args.workload, 2 if args.workload == "1080p" else 1)
            except Exception as e:  # the oracle binary did not travel / wrong ISA: report, do not fake
                cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
        nb = w * h * 3
        line = {"metric": "interpolated frames/sec (%s)" % args.model, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": ("f16 operands / f32 accumulate (split hi+lo operands in block heads%s)" % ("" if pmask == 15 else " and IFBlocks " + ",".join(str(k) for k in range(4) if not (pmask >> k) & 1)))
                         if args.precision == 1 else ("f32" if args.precision == 0 else "f16 / f32 accumulate"),
                "data": "synthetic",
                "config": {"workload": desc if args.model == MODEL else desc.replace("rife-v4.6", args.model), "timestep": args.timestep, "tta": args.tta,
                           "tta_temporal": args.tta_temporal, "pairs_per_step": PAIRS_PER_STEP, "precision_tier": args.precision, "lanes": args.lanes, "batch": args.batch, "plain_fp16_blocks_mask": eng.get_option("plain_blocks"), "recompute_fm": eng.get_option("recompute_fm"), "fused_v46_path": bool(eng.get_option("fast_active")),
                           "l2": "flushed between timed steps (256 MiB memset)", "weights": "reference model files" if "_ref" in md else "synthetic"},
                "gflop_per_frame": GFLOP_PER_FRAME[args.workload] if args.model == MODEL and not (args.tta or args.tta_temporal) else None,
                "model_tflops": value * GFLOP_PER_FRAME[args.workload] / 1000.0 if args.model == MODEL and not (args.tta or args.tta_temporal) else None,
                "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h_step),
                        "note": "bytes counted by the library; a frame shared by consecutive pairs of a batch is uploaded once"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "out_checksum": checksum}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
