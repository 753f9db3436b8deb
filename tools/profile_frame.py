"""Profiling target: a few 1080p frames through the fused rife-v4.6 path (one lane)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import parity
pkg = g.load_package()
w, h = (1920, 1080) if len(sys.argv) < 2 or sys.argv[1] == "1080p" else (3840, 2160)
a, b = parity.synth.pair(w, h)
r = pkg.RIFE(0, False, False, False, 1, False, True)
r.load(parity.model_dir("rife-v4.6"))
r.set_option("lanes", 1)
for _ in range(3):
    out = r.process(a, b, 0.5)
print("fast_active", r.get_option("fast_active"), int(out.sum()))
