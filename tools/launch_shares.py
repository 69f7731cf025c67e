#!/usr/bin/env python3
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: share of device time."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 8]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void pkv::<unnamed>::", "").replace("void pkv::", "")
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1000.0
        elif r[ui] == "ms":
            v *= 1000.0
        agg[name] = agg.get(name, 0.0) + v
        cnt[name] += 1
    tot = sum(agg.values())
    print(f"{path}: {sum(cnt.values())} launches, {tot / 1000.0:.3f} ms of device time (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda x: -x[1]):
        print(f"  {100 * v / tot:5.1f}%  {v / 1000.0:9.3f} ms  x{cnt[k]:4d}  {k}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
