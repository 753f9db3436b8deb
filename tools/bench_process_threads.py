"""Throughput of the reference-shaped call -- RIFE::process(in0, in1, t, out), one frame pair per call, host buffers --
when T threads call it on ONE handle, as the reference CLI's proc threads do (src/main.cpp:346-366).  Compares option
"combine" off (calls serialised, one pair at a time) and on (concurrent calls run as one lock-step batch)."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import __graft_entry__ as g
import parity

pkg = g.load_package()
w, h = (1920, 1080) if len(sys.argv) < 2 or sys.argv[1] == "1080p" else (3840, 2160)
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 24
frames = [torch.from_numpy(parity.synth.frame(k, w, h)).pin_memory() for k in range(9)]
r = pkg.RIFE(0, False, False, False, 1, False, True)
r.load(parity.model_dir("rife-v4.6"))
L = pkg.lib()
for combine in (0, 1):
    r.set_option("combine", combine)
    for T in (1, 2, 4, 8, 16):
        outs = [torch.empty_like(frames[0]).pin_memory() for _ in range(T)]
        gate = threading.Barrier(T + 1)

        def work(i):
            a, b, o = frames[i % 8].data_ptr(), frames[i % 8 + 1].data_ptr(), outs[i].data_ptr()
            L.rife_b200_process(r._h, a, b, w, h, 0.5, o)  # warm
            gate.wait()
            for _ in range(calls):
                L.rife_b200_process(r._h, a, b, w, h, 0.5, o)
            gate.wait()

        ts = [threading.Thread(target=work, args=(i,)) for i in range(T)]
        for t in ts:
            t.start()
        gate.wait()
        t0 = time.perf_counter()
        gate.wait()
        dt = time.perf_counter() - t0
        for t in ts:
            t.join()
        print("combine=%d threads=%2d: %7.1f process() calls/s (%dx%d, %d calls per thread)" % (combine, T, T * calls / dt, w, h, calls), flush=True)
print("combined batches %d for %d requests" % (r.get_option("combined_batches"), r.get_option("combined_requests")))
r.close()
