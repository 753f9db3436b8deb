#!/usr/bin/env python3
"""Parity of the fused path vs the oracle at 1080p for precision choices of the block-head tensors (option plain_blocks
bits 4-7: block k's head tensor / conv0 input read as plain fp16 instead of split hi+lo).  One oracle frame, several masks."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import parity  # noqa: E402
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
cases = {"synth": parity.synth.pair(w, h), "large_motion": parity.synth.pair(w, h, dx=24, dy=16)}
rows = []
for name, (a, b) in cases.items():
    ref, _ = parity.run_oracle("rife-v4.6", a, b, 0.5)
    for mask in (12, 12 | 0x80, 12 | 0xC0, 12 | 0xF0, 15, 15 | 0xF0):
        out = parity.run_gpu(pkg, "rife-v4.6", a, b, 0.5, options={"plain_blocks": mask})
        r = parity.compare(out, ref)
        r.update({"case": name, "mask": mask})
        rows.append(r)
        print(json.dumps(r))
