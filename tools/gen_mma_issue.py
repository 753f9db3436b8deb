#!/usr/bin/env python3
"""Generates rife-ncnn-vulkan_b200/csrc/tc_mma_issue.inc: one asm block per filter tap that issues all MT x NP MMAs of
the tap (MT accumulators x NP operand planes share the tap's B matrix).  WS = 1 emits the weight-stationary form
(tcgen05.mma.ws, collector b0: the first MMA fills the B collector, the following ones reuse it)."""
import os

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rife-ncnn-vulkan_b200", "csrc", "tc_mma_issue.inc")


def block(mt, np_, ws):
    lines = ['"{\\n"', '".reg .pred q, p, pt;\\n"', '".reg .b64 da, db;\\n"', '".reg .b32 al, tm;\\n"', '"elect.sync _|q, 0xffffffff;\\n"',
             '"setp.ne.b32 p, %5, 0;\\n"', '"setp.eq.b32 pt, 0, 0;\\n"', '"mov.b64 db, {%2, %3};\\n"']
    ops = []
    n = mt * np_
    k = 0
    for m in range(mt):
        for pl in range(np_):
            ia, it = 6 + 2 * k, 7 + 2 * k
            lines += ['"add.u32 al, %%1, %%%d;\\n"' % ia, '"add.u32 tm, %%0, %%%d;\\n"' % it, '"mov.b64 da, {al, %3};\\n"']
            pred = "p" if pl == 0 else "pt"
            if ws:
                coll = "fill" if k == 0 else ("lastuse" if k == n - 1 else "use")
                if n == 1:
                    coll = "discard"
                lines.append('"@q tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::%s [tm], da, db, %%4, %s;\\n"' % (coll, pred))
            else:
                lines.append('"@q tcgen05.mma.cta_group::1.kind::f16 [tm], da, db, %%4, %s;\\n"' % pred)
            ops += ['"n"(%d * APLANE16 + %d * ROWSTEP16)' % (pl, m), '"n"(%d * NCOLS)' % m]
            k += 1
    lines.append('"}\\n"')
    body = "\n            ".join(lines)
    return ("        asm volatile(\n            %s\n            ::\"r\"(acc), \"r\"(a_lo), \"r\"(b_lo), \"r\"(desc_hi), \"r\"(idesc), \"r\"(accumulate), %s);\n"
            % (body, ", ".join(ops)))


def pair_block(mt, np_):
    """Even halo-row views of one kernel column dx (2-row accumulators): view j = rows 2j, 2j+1 is the dy=0 operand of
    accumulator j and the dy=2 operand of accumulator j-1, so for 1 <= j <= mt-1 ONE MMA of 2N columns updates both
    ([D_{j-1} | D_j] += A_j * [W_dy2 | W_dy0]); j = 0 and j = mt are the unpaired ends (N columns)."""
    lines = ['"{\\n"', '".reg .pred q, p, pt;\\n"', '".reg .b64 da, db, db0;\\n"', '".reg .b32 al, tm, bl;\\n"', '"elect.sync _|q, 0xffffffff;\\n"',
             '"setp.ne.b32 p, %6, 0;\\n"', '"setp.eq.b32 pt, 0, 0;\\n"', '"mov.b64 db, {%2, %3};\\n"', '"add.u32 bl, %2, %7;\\n"', '"mov.b64 db0, {bl, %3};\\n"']
    ops = ['"n"(NCOLS)']  # %7: rows (16-byte units) from the dy=2 block to the dy=0 block
    k = 0
    for j in range(mt + 1):
        for pl in range(np_):
            ia, it = 8 + 2 * k, 9 + 2 * k
            lines += ['"add.u32 al, %%1, %%%d;\\n"' % ia, '"add.u32 tm, %%0, %%%d;\\n"' % it, '"mov.b64 da, {al, %3};\\n"']
            pred = "p" if pl == 0 else "pt"
            if j == 0:
                bd, idc = "db0", "%4"      # dy = 0 weights only -> accumulator 0
            elif j == mt:
                bd, idc = "db", "%4"       # dy = 2 weights only -> accumulator mt-1
            else:
                bd, idc = "db", "%5"       # [dy2 | dy0] -> accumulators j-1, j
            lines.append('"@q tcgen05.mma.cta_group::1.kind::f16 [tm], da, %s, %s, %s;\\n"' % (bd, idc, pred))
            ops += ['"n"(%d * APLANE16 + %d * ROWSTEP16)' % (pl, j), '"n"(%d * NCOLS)' % max(j - 1, 0)]
            k += 1
    lines.append('"}\\n"')
    body = "\n            ".join(lines)
    return ("        asm volatile(\n            %s\n            ::\"r\"(acc), \"r\"(a_lo), \"r\"(b_lo), \"r\"(desc_hi), \"r\"(idesc), \"r\"(idesc2), \"r\"(accumulate), %s);\n"
            % (body, ", ".join(ops)))


def wide_block(mt, np_):
    """One-row accumulators (128 consecutive pixels of ONE image row each): halo row r is the dy operand of accumulator r - dy
    for dy = 0, 1, 2, i.e. of up to THREE accumulators that are adjacent in TMEM, so ONE MMA of up to 3N columns against
    [W_dy2 | W_dy1 | W_dy0] updates all of them: mt + 2 MMAs per kernel column instead of 3 * mt, the activation row is read
    from shared memory once instead of three times."""
    lines = ['"{\\n"', '".reg .pred q, p, pt;\\n"', '".reg .b64 da, db;\\n"', '".reg .b32 al, tm, bl;\\n"', '"elect.sync _|q, 0xffffffff;\\n"',
             '"setp.ne.b32 p, %7, 0;\\n"', '"setp.eq.b32 pt, 0, 0;\\n"']
    ops = []
    k = 0
    for r in range(mt + 2):
        dy_max, dy_min = min(2, r), max(0, r - (mt - 1))
        ncol = dy_max - dy_min + 1            # accumulators covered
        idc = {1: "%4", 2: "%5", 3: "%6"}[ncol]
        for pl in range(np_):
            ia, it, ib = 8 + 3 * k, 9 + 3 * k, 10 + 3 * k
            lines += ['"add.u32 al, %%1, %%%d;\\n"' % ia, '"add.u32 tm, %%0, %%%d;\\n"' % it, '"add.u32 bl, %%2, %%%d;\\n"' % ib, '"mov.b64 da, {al, %3};\\n"', '"mov.b64 db, {bl, %3};\\n"']
            pred = "p" if pl == 0 else "pt"
            lines.append('"@q tcgen05.mma.cta_group::1.kind::f16 [tm], da, db, %s, %s;\\n"' % (idc, pred))
            ops += ['"n"(%d * APLANE16 + %d * ROWSTEP16)' % (pl, r), '"n"(%d * NCOLS)' % (r - dy_max), '"n"(%d * NCOLS)' % (2 - dy_max)]
            k += 1
    lines.append('"}\\n"')
    body = "\n            ".join(lines)
    return ("        asm volatile(\n            %s\n            ::\"r\"(acc), \"r\"(a_lo), \"r\"(b_lo), \"r\"(desc_hi), \"r\"(idesc), \"r\"(idesc2), \"r\"(idesc3), \"r\"(accumulate), %s);\n"
            % (body, ", ".join(ops)))


def chunk_block(n, mt, np_):
    """ALL MMAs of one 16-channel chunk of a paired-layout 3x3 stride-1 conv in ONE asm block (27 + MT taps x NP planes): per kernel
    column dx the paired rows, then the dy = 1 taps, finally the narrow identity tap (predicated on %6).  Same order as the per-tap
    blocks it replaces (bit-identical accumulation); every operand offset is a literal, so between two MMAs there is nothing
    but three uniform adds -- the per-block preamble (elect, register -> uniform-register moves, divergence checks: ~40 SASS
    instructions) is paid once per chunk instead of seven times."""
    twp, rowstep = 64, 128
    aplane = 2 * (2 * mt + 2) * twp  # 16-byte units
    L = ['"{\\n"', '".reg .pred q, pi, qi, pt;\\n"', '"setp.eq.b32 pt, 0, 0;\\n"', '".reg .b64 da, db;\\n"', '".reg .b32 al, tm, bl;\\n"', '"elect.sync _|q, 0xffffffff;\\n"',
         '"setp.ne.b32 pi, %6, 0;\\n"', '"and.pred qi, q, pi;\\n"']

    def mma(a_off, b_reg, b_off, t_reg, t_off, idc, pred="q"):
        return ['"add.u32 al, %%1, %d;\\n"' % a_off, '"add.u32 bl, %s, %d;\\n"' % (b_reg, b_off), '"add.u32 tm, %s, %d;\\n"' % (t_reg, t_off),
                '"mov.b64 da, {al, %3};\\n"', '"mov.b64 db, {bl, %3};\\n"', '"@%s tcgen05.mma.cta_group::1.kind::f16 [tm], da, db, %s, pt;\\n"' % (pred, idc)]
    for dx in range(3):
        bblk = dx * 2 * 3 * n
        for j in range(mt + 1):
            for pl in range(np_):
                a_off = dx + j * rowstep + pl * aplane
                if j == 0:
                    L += mma(a_off, "%2", bblk + n, "%0", 0, "%4")
                elif j == mt:
                    L += mma(a_off, "%2", bblk, "%0", (mt - 1) * n, "%4")
                else:
                    L += mma(a_off, "%2", bblk, "%0", (j - 1) * n, "%5")
        for m in range(mt):
            for pl in range(np_):
                L += mma(dx + twp + m * rowstep + pl * aplane, "%2", bblk + 2 * n, "%0", m * n, "%4")
    for m in range(mt):
        for pl in range(np_):
            L += mma(twp + 1 + m * rowstep + pl * aplane, "%7", 0, "%9", m * n, "%8", pred="qi")
    L.append('"}\\n"')
    body = "\n            ".join(L)
    return ("        asm volatile(\n            %s\n            ::\"r\"(acc), \"r\"(a_lo), \"r\"(b_lo), \"r\"(desc_hi), \"r\"(idesc), \"r\"(idesc2), \"r\"(ident_on), \"r\"(ident_b_lo), \"r\"(idesc16), \"r\"(acc_ident));\n" % body)


def main():
    o = ["// Generated by tools/gen_mma_issue.py -- do not edit.  One asm block per filter tap issues all MT x NP MMAs of that tap;",
         "// descriptor / TMEM address arithmetic stays inside the block so ptxas keeps it on the uniform datapath.",
         "// WS = 1: weight-stationary form -- the tap's B matrix is read from shared memory once (collector fill) and reused by",
         "// the following MMAs of the block instead of being fetched again for every accumulator.",
         "template <int MT, int NP, int APLANE16, int ROWSTEP16, int NCOLS, int WS = 0>",
         "__device__ __forceinline__ void umma_issue_tap(uint32_t acc, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {"]
    first = True
    for ws in (0, 1):
        for mt in (1, 2, 4):
            for np_ in (1, 2):
                o.append("    %sif constexpr (MT == %d && NP == %d && WS == %d) {" % ("" if first else "else ", mt, np_, ws))
                o.append(block(mt, np_, ws).rstrip("\n"))
                o.append("    }")
                first = False
    o.append("}")
    o += ["// Paired issue (see tools/gen_mma_issue.py pair_block): the even halo-row views of one kernel column.  b_lo addresses the",
          "// [dy2 | dy0 | dy1] weight block of that column (rows of 16 B; dy0 starts NCOLS rows in); idesc = N columns, idesc2 = 2N.",
          "template <int MT, int NP, int APLANE16, int ROWSTEP16, int NCOLS>",
          "__device__ __forceinline__ void umma_issue_pair(uint32_t acc, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t idesc2, uint32_t accumulate) {"]
    first = True
    for mt in (1, 2, 4):
        for np_ in (1, 2):
            o.append("    %sif constexpr (MT == %d && NP == %d) {" % ("" if first else "else ", mt, np_))
            o.append(pair_block(mt, np_).rstrip("\n"))
            o.append("    }")
            first = False
    o.append("}")
    o += ["// Wide issue (see tools/gen_mma_issue.py wide_block): one-row accumulators, the halo rows of one kernel column.  b_lo addresses",
          "// the [dy2 | dy1 | dy0] weight block of that column (rows of 16 B); idesc / idesc2 / idesc3 = N / 2N / 3N columns.",
          "template <int MT, int NP, int APLANE16, int ROWSTEP16, int NCOLS>",
          "__device__ __forceinline__ void umma_issue_wide(uint32_t acc, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t idesc2, uint32_t idesc3, uint32_t accumulate) {"]
    first = True
    for mt in (2, 4):
        for np_ in (1, 2):
            o.append("    %sif constexpr (MT == %d && NP == %d) {" % ("" if first else "else ", mt, np_))
            o.append(wide_block(mt, np_).rstrip("\n"))
            o.append("    }")
            first = False
    o.append("}")
    o += ["// Whole-chunk issue (see tools/gen_mma_issue.py chunk_block): every MMA of one 16-channel chunk of a paired 3x3 stride-1 conv in one",
          "// asm block with literal operand offsets.  a_lo = slab address >> 4 | LBO, b_lo = address of the chunk's [dx][half][3N] weight blocks >> 4 | LBO(3N rows).",
          "template <int N, int MT, int NP>",
          "__device__ __forceinline__ void umma_issue_chunk_paired(uint32_t acc, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t idesc2, uint32_t ident_on,",
          "                                                        uint32_t ident_b_lo, uint32_t idesc16, uint32_t acc_ident) {"]
    first = True
    for n, mt in ((16, 4), (32, 4), (48, 4), (64, 4), (96, 2), (128, 2)):
        for np_ in (1, 2):
            o.append("    %sif constexpr (N == %d && MT == %d && NP == %d) {" % ("" if first else "else ", n, mt, np_))
            o.append(chunk_block(n, mt, np_).rstrip("\n"))
            o.append("    }")
            first = False
    o.append("}")
    open(OUT, "w").write("\n".join(o) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
