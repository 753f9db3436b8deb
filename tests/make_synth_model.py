#!/usr/bin/env python3
"""Writes a model directory with the rife-v4.6 IFNet architecture and seeded random weights in
the reference's on-disk format (ncnn .param text + .bin with fp16 conv weights).

Used only when the reference's model files did not travel (no oracle/_ref/models): bench.py then reports
`"weights": "synthetic"` and the parity tests compare the CUDA path with the oracle restatement on this model.
The graph is written from the architecture description (blocks of conv3x3 s2, conv3x3 s2, 8 x residual conv3x3,
deconv4x4 s2 + PixelShuffle; bilinear resampling, warps, flow / mask accumulation, sigmoid blend); blob and layer
names are ours except the interface blobs in0, in1, in2, flow0..flow3, out0 that the engine addresses by name."""
import os
import struct
import sys

import numpy as np


class Graph:
    def __init__(self):
        self.layers = []   # (type, name, bottoms, tops, params-string)
        self.weights = []  # per layer: list of ("fp16"|"f32", ndarray)
        self.n = 0

    def uid(self, prefix):
        self.n += 1
        return "%s_%d" % (prefix, self.n)

    def add(self, typ, bottoms, ntop=1, params="", weights=None, top_names=None, name=None):
        name = name or self.uid(typ.lower().replace(".", "_"))
        tops = top_names or [self.uid("b") for _ in range(ntop)]
        self.layers.append([typ, name, list(bottoms), tops, params])
        self.weights.append(weights or [])
        return tops[0] if ntop == 1 else tops

    def finalize(self):
        """Insert ncnn-style Split layers for blobs with several consumers."""
        consumers = {}
        for li, L in enumerate(self.layers):
            for bi, b in enumerate(L[2]):
                consumers.setdefault(b, []).append((li, bi))
        out_layers, out_weights = [], []
        rename = {}  # (layer index, bottom index) -> new blob name
        for li, L in enumerate(self.layers):
            bottoms = [rename.get((li, bi), b) for bi, b in enumerate(L[2])]
            out_layers.append([L[0], L[1], bottoms, L[3], L[4]])
            out_weights.append(self.weights[li])
            for t in L[3]:
                cs = consumers.get(t, [])
                if len(cs) > 1:
                    names = ["%s_s%d" % (t, k) for k in range(len(cs))]
                    out_layers.append(["Split", "split_" + t, [t], names, ""])
                    out_weights.append([])
                    for (cl, cb), nn in zip(cs, names):
                        rename[(cl, cb)] = nn
        blobs = set()
        for L in out_layers:
            blobs.update(L[2])
            blobs.update(L[3])
        return out_layers, out_weights, len(blobs)


def build_v46(seed=0):
    rng = np.random.default_rng(seed)
    g = Graph()

    def conv(x, cin, cout, stride, act_leaky):
        w = (rng.standard_normal((cout, cin, 9)) * np.sqrt(2.0 / (9 * cin)) * 0.7).astype(np.float16)
        b = (rng.standard_normal(cout) * 0.02).astype(np.float32)
        p = "0=%d 1=3 %s4=1 5=1 6=%d" % (cout, "3=2 " if stride == 2 else "", w.size)
        if act_leaky:
            p += " 9=2 -23310=1,2.000000e-01"
        return g.add("Convolution", [x], params=p, weights=[("fp16", w), ("f32", b)])

    def deconv(x, cin):
        w = (rng.standard_normal((24, cin, 16)) * np.sqrt(1.0 / (4 * cin)) * 0.15).astype(np.float16)
        b = (rng.standard_normal(24) * 0.01).astype(np.float32)
        return g.add("Deconvolution", [x], params="0=24 1=4 3=2 4=1 5=1 6=%d" % w.size, weights=[("fp16", w), ("f32", b)])

    def interp(x, s):
        return g.add("Interp", [x], params="0=2 1=%e 2=%e" % (s, s))

    def crop(x, a, b):
        return g.add("Crop", [x], params="-23309=1,%d -23310=1,%d -23311=1,0" % (a, b))

    in0 = g.add("Input", [], top_names=["in0"], name="in0")
    in1 = g.add("Input", [], top_names=["in1"], name="in1")
    in2 = g.add("Input", [], top_names=["in2"], name="in2")
    widths, scales = [192, 128, 96, 64], [8, 4, 2, 1]
    F = M = None
    for k in range(4):
        c, s = widths[k], scales[k]
        if k == 0:
            x = interp(g.add("Concat", [in0, in1, in2]), 1.0 / s)
            cin = 7
        else:
            w1 = g.add("rife.Warp", [in1, crop(F, 2, 4)])
            w0 = g.add("rife.Warp", [in0, crop(F, 0, 2)])
            x = g.add("Concat", [w0, w1, in2, M])
            if s != 1:
                x = interp(x, 1.0 / s)
                fd = g.add("BinaryOp", [interp(F, 1.0 / s)], params="0=3 1=1 2=%e" % float(s))
            else:
                fd = F
            x = g.add("Concat", [x, fd])
            cin = 12
        y = conv(x, cin, c // 2, 2, True)
        y = conv(y, c // 2, c, 2, True)
        for _ in range(8):
            t = conv(y, c, c, 1, False)
            t = g.add("BinaryOp", [t, y], params="")
            y = g.add("ReLU", [t], params="0=2.000000e-01")
        d = g.add("PixelShuffle", [deconv(y, c)], params="0=2", top_names=["flow%d" % k])
        u = interp(d, float(s)) if s != 1 else d
        uf, um = crop(u, 0, 4), crop(u, 4, 5)
        if k == 0:
            F = g.add("BinaryOp", [uf], params="0=2 1=1 2=%e" % float(s))
            M = um
        elif s != 1:
            F = g.add("Eltwise", [F, uf], params="0=1 -23301=2,1.000000e+00,%e" % float(s))
            M = g.add("BinaryOp", [M, um], params="")
        else:
            F = g.add("BinaryOp", [F, uf], params="")
            M = g.add("BinaryOp", [M, um], params="")
    m = g.add("Sigmoid", [M])
    om = g.add("BinaryOp", [m], params="0=7 1=1 2=1.000000e+00")
    t1 = g.add("BinaryOp", [g.add("rife.Warp", [in1, crop(F, 2, 4)]), om], params="0=2")
    t0 = g.add("BinaryOp", [g.add("rife.Warp", [in0, crop(F, 0, 2)]), m], params="0=2")
    g.add("BinaryOp", [t0, t1], params="", top_names=["out0"])
    return g


def write_model(dirpath, seed=0):
    os.makedirs(dirpath, exist_ok=True)
    layers, weights, nblobs = build_v46(seed).finalize()
    with open(os.path.join(dirpath, "flownet.param"), "w") as f:
        f.write("7767517\n%d %d\n" % (len(layers), nblobs))
        for typ, name, bottoms, tops, params in layers:
            f.write("%-24s %-24s %d %d %s %s\n" % (typ, name, len(bottoms), len(tops), " ".join(bottoms + tops), params))
    with open(os.path.join(dirpath, "flownet.bin"), "wb") as f:
        for ws in weights:
            for kind, arr in ws:
                if kind == "fp16":
                    f.write(struct.pack("<I", 0x01306B47))
                    raw = np.ascontiguousarray(arr, dtype=np.float16).tobytes()
                    f.write(raw)
                    f.write(b"\0" * ((-len(raw)) % 4))
                else:
                    f.write(np.ascontiguousarray(arr, dtype=np.float32).tobytes())
    return dirpath


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "models", "rife-v4.6")
    print(write_model(out))
