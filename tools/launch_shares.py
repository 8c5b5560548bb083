#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel counts,
total and share of device time.  usage: launch_shares.py launches.csv out.txt"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# %s: per-launch device times under ncu (cold cache, serialised: compare shares)\n" % src.split("/")[-1])
        f.write("%-64s %7s %12s %10s %10s %7s\n" % ("kernel", "launches", "total_us", "min_us", "max_us", "share"))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-64s %7d %12.1f %10.2f %10.2f %6.1f%%\n" % (k[:64], a[0], a[1], a[2], a[3], 100 * a[1] / tot))
    print(open(out).read())


if __name__ == "__main__":
    main()
