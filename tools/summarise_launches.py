#!/usr/bin/env python3
"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> per-kernel totals (count, total us, share)."""
import collections
import csv
import re
import sys


def main(path, top=40):
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines[start:]):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit.startswith("ns") else (v if unit.startswith("us") else v * 1000.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("%d launches, %.1f us in total" % (sum(a[0] for a in agg.values()), tot))
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-90s %6d %12.1f us %5.1f %%" % (name[:90], n, us, 100 * us / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
