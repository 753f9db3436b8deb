"""Epilogue knock-out experiment on the tcgen05 conv (timing only): which part of the epilogue sets the tile period?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

pkg = g.load_package()
names = {0: "full", 8: "empty epilogue"}
for split in (False, True):
    for flags in (0, 8):
        buf = pkg.debug_conv_timeline(64, 64, 272, 480, split=split, batch=8, skip_tiles=5, flags=flags).astype(np.int64)
        per, epi, tot = [], [], []
        for cta in range(0, 148, 7):
            r = buf[cta]
            e = r[44:52]
            per += [int(e[2 * i + 2] - e[2 * i]) for i in range(3)]
            epi += [int(e[2 * i + 1] - e[2 * i]) for i in range(4)]
            tot.append(int(r[56] - r[0]))
        print("split=%d %-20s tile period %6.0f  epilogue %6.0f  kernel cycles %7.0f" % (split, names[flags], np.median(per), np.median(epi), np.median(tot)))
