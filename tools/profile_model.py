#!/usr/bin/env python3
"""Runs a few frames of one model / mode through rife_b200_process_device (frames resident in HBM) -- the body of an
`ncu --metrics gpu__time_duration.sum` launch list for the configurations bench.py does not headline (e.g. BASELINE configs[3]:
rife-anime 1080p -x -z), and a plain fps print-out without ncu."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="rife-anime")
    ap.add_argument("--w", type=int, default=1920)
    ap.add_argument("--h", type=int, default=1080)
    ap.add_argument("--tta", action="store_true")
    ap.add_argument("--tta-temporal", action="store_true")
    ap.add_argument("--uhd", action="store_true")
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--precision", type=int, default=1)
    ap.add_argument("--lanes", type=int, default=0, help="0 = library default (2; 8 for spatial-TTA engines, whose orientations are dealt to the lanes)")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    import parity
    pkg = g.load_package()
    v2, v4 = pkg.family_flags(args.model)
    eng = pkg.RIFE(0, args.tta, args.tta_temporal, args.uhd, 1, v2, v4)
    eng.load(parity.model_dir(args.model))
    eng.set_option("precision", args.precision)
    if args.lanes > 0:
        eng.set_option("lanes", args.lanes)
    a, b = parity.synth.pair(args.w, args.h)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty_like(da)
    for _ in range(args.warmup):
        eng.process_ptr(da.data_ptr(), db.data_ptr(), args.w, args.h, 0.5, out.data_ptr(), device=True)
    torch.cuda.synchronize()
    l0 = pkg.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.frames):
        eng.process_ptr(da.data_ptr(), db.data_ptr(), args.w, args.h, 0.5, out.data_ptr(), device=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s %dx%d tta=%d tta_temporal=%d precision=%d lanes=%d: %.3f frames/s (%.1f ms/frame, %d launches/frame)" %
          (args.model, args.w, args.h, args.tta, args.tta_temporal, args.precision, eng.get_option("lanes"), args.frames / dt, 1000 * dt / args.frames, (pkg.launch_count() - l0) // args.frames))


if __name__ == "__main__":
    main()
