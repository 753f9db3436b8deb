#!/usr/bin/env python3
"""Rate of device -> host copies (copy engine, pinned destination) while the fused rife-v4.6 path keeps the GPU busy on other
streams, against the same copies on an idle GPU: what bounds the end-to-end legs of bench.py."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as g
import parity

pkg = g.load_package()
w, h = 1920, 1080
eng = pkg.RIFE(0, False, False, False, 1, False, True)
eng.load(parity.model_dir("rife-v4.6"))
frames = [torch.from_numpy(parity.synth.frame(k, w, h)).cuda() for k in range(9)]
outs = [torch.empty_like(frames[0]) for _ in range(64)]
d0 = [frames[i % 8].data_ptr() for i in range(64)]
d1 = [frames[i % 8 + 1].data_ptr() for i in range(64)]
do = [o.data_ptr() for o in outs]
stop = False

def load():
    while not stop:
        eng.process_batch_ptr(d0, d1, w, h, [0.5] * 64, do, device=True)

n = 64 << 20
src = torch.empty(n, dtype=torch.uint8, device="cuda")
dst = torch.empty(n, dtype=torch.uint8).pin_memory()
hsrc = torch.empty(n, dtype=torch.uint8).pin_memory()
ddst = torch.empty(n, dtype=torch.uint8, device="cuda")
cs = torch.cuda.Stream()

def rate(a, b, reps=8):
    best, tot = None, 0.0
    with torch.cuda.stream(cs):
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cs); a.copy_(b, non_blocking=True); e1.record(cs); cs.synchronize()
            ms = e0.elapsed_time(e1); tot += ms
            best = ms if best is None or ms < best else best
    return n / (best * 1e-3) / 1e9, n * reps / (tot * 1e-3) / 1e9

print("idle GPU : D2H best %.1f GB/s, mean %.1f | H2D best %.1f, mean %.1f" % (rate(dst, src) + rate(ddst, hsrc)))
t = threading.Thread(target=load); t.start(); time.sleep(0.5)
print("under the rife-v4.6 step: D2H best %.1f GB/s, mean %.1f | H2D best %.1f, mean %.1f" % (rate(dst, src, 16) + rate(ddst, hsrc, 16)))
stop = True; t.join()
