"""Profiling target: the dominant tcgen05 conv (64->64 at quarter resolution of 1080p / 4K) launched alone."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
h, w = (272, 480) if len(sys.argv) < 2 or sys.argv[1] == "1080p" else (544, 960)
split = int(sys.argv[2]) if len(sys.argv) > 2 else 1
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
s = torch.cuda.Stream()
pkg.bench_conv(s.cuda_stream, 64, 64, h, w, split, 4, batch=batch)
torch.cuda.synchronize()
