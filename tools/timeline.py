"""Prints the in-kernel clock64 timeline of the tcgen05 conv (steady state, batched) -- diagnostics for profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

pkg = g.load_package()
cases = [(64, 480, 272, False, 8, 5), (64, 480, 272, True, 8, 5), (64, 480, 272, False, 1, 0), (96, 240, 136, False, 8, 3), (64, 960, 544, False, 2, 5)]
for c, w, h, split, batch, skip in cases:
    buf = pkg.debug_conv_timeline(c, c, h, w, split=split, batch=batch, skip_tiles=skip).astype(np.int64)
    for cta in (0, 77):
        r = buf[cta]
        t0 = r[0]
        rel = lambda a: [int(v - t0) if v else -1 for v in a]
        full, comm, epi = rel(r[16:28]), rel(r[32:44]), rel(r[44:52])
        print("c=%d %dx%d split=%d batch=%d skip=%d cta %d: end %d" % (c, w, h, split, batch, skip, cta, r[56] - t0))
        print("   prod issue   ", rel(r[1:13]))
        print("   mma full seen", full)
        print("   mma committed", comm)
        print("   mma stage dur", [b - a for a, b in zip(full, comm)])
        print("   epi [full,done]x4", epi, "epi dur", [epi[2 * i + 1] - epi[2 * i] for i in range(4)], "tile period", [epi[2 * i + 2] - epi[2 * i] for i in range(3)])
