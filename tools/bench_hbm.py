#!/usr/bin/env python3
"""Achieved HBM bandwidth of the HBM-side kernels (csrc/hbm_kernels.cu) at 1080p / 4K against the measured copy peak
(MEASURED_PEAKS.json), CUDA-event timed on the launching stream, L2 flushed between samples.  Bytes = inputs read once +
outputs written once at their storage type.  `--ncu` makes it a short run for an ncu capture."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1080p", choices=["1080p", "4k"])
    ap.add_argument("--ncu", action="store_true")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    pkg = g.load_package()
    w, h = (1920, 1080) if args.size == "1080p" else (3840, 2160)
    wp, hp = (w + 31) // 32 * 32, (h + 31) // 32 * 32
    plane, n = wp * hp, w * h
    fw, fh = wp // 2, hp // 2  # v1/v2 flow resolution
    cases = [  # name, which, (w, h, c), algorithmic bytes per launch
        ("preproc x8 orientations", "preproc", (w, h, 8), n * 3 + 8 * 3 * plane * 4),
        ("preproc x1", "preproc", (w, h, 1), n * 3 + 3 * plane * 4),
        ("postproc plain", "postproc", (w, h, 1), 3 * n * 4 + n * 3),
        ("postproc temporal (2 in)", "postproc", (w, h, 2), 2 * 3 * n * 4 + n * 3),
        ("postproc tta (8 in)", "postproc", (w, h, 8), 8 * 3 * plane * 4 + n * 3),
        ("postproc tta+temporal (16 in)", "postproc", (w, h, 16), 16 * 3 * plane * 4 + n * 3),
        ("flow_tta_avg 5ch (v4, full res)", "flow_tta_avg", (wp, hp, 5), 2 * 8 * 5 * plane * 4),
        ("flow_tta_avg 4ch (v2, half res)", "flow_tta_avg", (fw, fh, 4), 2 * 8 * 4 * fw * fh * 4),
        ("temporal_merge_v2 +mask", "temporal_merge_v2", (wp, hp, 1), 4 * 5 * plane * 4),
        ("warp 3ch (frame)", "warp", (wp, hp, 3), (3 + 2 + 3) * plane * 4),
        ("warp 32ch (context, half res)", "warp", (fw, fh, 32), (32 + 2 + 32) * fw * fh * 4),
    ]
    peak = 6572.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    st = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for name, which, (cw, chh, cc), nbytes in cases:
        with torch.cuda.stream(st):
            pkg.debug_hbm(which, cw, chh, cc, iters=2, cuda_stream_ptr=st.cuda_stream)
            torch.cuda.synchronize()
            if args.ncu:
                continue
            best = None
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                pkg.debug_hbm(which, cw, chh, cc, iters=1, cuda_stream_ptr=st.cuda_stream)
                e1.record(st)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None or ms < best else best
        gbs = nbytes / (best * 1e-3) / 1e9
        rows.append({"kernel": name, "us": round(best * 1000, 1), "MB": round(nbytes / 1e6, 1), "GB_s": round(gbs, 1), "frac_of_copy_peak": round(gbs / peak, 3)})
        print("%-34s %8.1f us %8.1f MB %8.1f GB/s  %.2f of the %.0f GB/s copy peak" % (name, best * 1000, nbytes / 1e6, gbs, gbs / peak, peak))
    if not args.ncu:
        print(json.dumps({"size": args.size, "hbm_peak_gbs": peak, "l2": "flushed before every sample (256 MiB memset)", "rows": rows}))


if __name__ == "__main__":
    main()
