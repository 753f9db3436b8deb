"""Duration of the tcgen05 conv (CUDA events, no in-kernel instrumentation) with parts of the mainloop knocked out
(TcConvArgs::dbg_flags via RIFE_B200_DBG_FLAGS; results are wrong, timing only): how much of a launch is MMA execution and how
much is per-stage / per-tile fixed cost.  flags: 1 no identity tap, 2 no bias MMAs, 4 a third of the 3x3 taps, 8 empty epilogue, 16 no loads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g

pkg = g.load_package()
NAMES = {1: "no identity tap", 2: "no bias MMAs", 4: "1/3 of the taps", 8: "empty epilogue", 16: "no loads"}
st = torch.cuda.Stream()
for c, w, h, batch in [(64, 480, 272, 8), (96, 240, 136, 8), (128, 120, 68, 8), (192, 60, 34, 8)]:
    for split in (0, 1):
        row = []
        for fl in (0, 1, 2, 4, 8, 16, 15, 23, 24, 31):
            os.environ["RIFE_B200_DBG_FLAGS"] = str(fl)
            with torch.cuda.stream(st):
                pkg.bench_conv(st.cuda_stream, c, c, h, w, split, 3, batch=batch)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                pkg.bench_conv(st.cuda_stream, c, c, h, w, split, 20, batch=batch)
                e1.record(st)
                torch.cuda.synchronize()
            row.append((fl, e0.elapsed_time(e1) / 20 * 1000))
        print("N=%3d %dx%d x%d split=%d:" % (c, w, h, batch, split), "  ".join("[%s] %.1f us" % ("+".join(NAMES[b] for b in (1, 2, 4, 8, 16) if fl & b) or "full", us) for fl, us in row))
os.environ.pop("RIFE_B200_DBG_FLAGS", None)
