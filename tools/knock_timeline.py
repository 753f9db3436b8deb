"""In-kernel clock64 timeline of the 64->64 conv (8 x 480x272, plain fp16) under the knock-out flags of tools/knockout.py:
tile period seen by the epilogue warps, MMA-warp issue time per 16-channel chunk, gap between chunks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

pkg = g.load_package()
NAMES = {1: "no identity tap", 2: "no bias MMAs", 4: "1/3 taps", 8: "empty epilogue", 16: "no loads"}
for c, w, h, batch in [(64, 480, 272, 8), (96, 240, 136, 8)]:
    for fl in (0, 4, 7, 8, 16, 15, 23, 31):
        buf = pkg.debug_conv_timeline(c, c, h, w, split=False, batch=batch, skip_tiles=4, flags=fl).astype(np.int64)
        per, issue, gap, total = [], [], [], []
        for cta in range(0, 148, 3):
            r = buf[cta]
            epi = r[44:52]
            per += [int(epi[2 * i + 2] - epi[2 * i]) for i in range(3) if epi[2 * i + 2] > epi[2 * i] > 0]
            full, comm = r[16:28], r[32:44]
            issue += [int(comm[i] - full[i]) for i in range(11) if comm[i] > full[i] > 0]
            gap += [int(full[i + 1] - comm[i]) for i in range(11) if full[i + 1] > comm[i] > 0]
            if r[56] > r[0] > 0:
                total.append(int(r[56] - r[0]))
        desc = " + ".join(NAMES[b] for b in (1, 2, 4, 8, 16) if fl & b) or "full"
        print("N=%3d flags %2d (%-58s): tile period %6.0f | chunk issue %5.0f, gap %5.0f | kernel %7.0f cycles (medians)" %
              (c, fl, desc, np.median(per) if per else -1, np.median(issue) if issue else -1, np.median(gap) if gap else -1, np.median(total) if total else -1))
