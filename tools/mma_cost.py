"""Cycles per tcgen05.mma as a function of N (128 x N x 16, both operands in shared memory): steady-state stage
durations from the in-kernel clock64 timeline with an empty epilogue, so only the mainloop is measured."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

pkg = g.load_package()
MT = {32: 4, 48: 4, 64: 4, 96: 2, 128: 2, 192: 1}
for c, w, h, batch in [(32, 480, 272, 8), (48, 480, 272, 8), (64, 480, 272, 8), (96, 480, 272, 8), (128, 480, 272, 8), (192, 480, 272, 8)]:
    for split, fl in ((False, 8), (False, 72), (True, 8), (True, 72)):
        buf = pkg.debug_conv_timeline(c, c, h, w, split=split, batch=batch, skip_tiles=4, flags=fl).astype(np.int64)
        per = []
        for cta in range(0, 148, 5):
            r = buf[cta]
            full, comm = r[16:28], r[32:44]
            # committed[i] -> committed[i+1] = one pipeline stage of the issue loop in steady state
            per += [int(comm[i + 1] - comm[i]) for i in range(10) if comm[i + 1] > comm[i] > 0]
        mmas = (9 + (1 if c <= 128 else 0)) * MT[c] * (2 if split else 1)  # taps (+ identity tap) x accumulators x planes per 16-channel stage
        print("N=%3d split=%d%s: stage %6.0f cycles, %3d MMAs -> %5.1f cycles per MMA (operand bytes A 4096 + B %d)" % (c, split, " +commit" if fl & 64 else "", np.median(per), mmas, np.median(per) / mmas, c * 32))
