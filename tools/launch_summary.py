#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches and mean / total time."""
import collections
import csv
import sys


def main(path, last=None):
    rows = list(csv.reader(open(path)))
    hdr = None; data = []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r; continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    data = [d for d in data if d.get("Metric Name") == "gpu__time_duration.sum"]
    if last:
        data = data[-int(last):]
    agg = collections.OrderedDict()
    for d in data:
        v = float(d["Metric Value"].replace(",", ""))
        if d["Metric Unit"] in ("ns", "nsecond"): v /= 1e3
        elif d["Metric Unit"] in ("ms", "msecond"): v *= 1e3
        agg.setdefault(d["Kernel Name"][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print("%-72s n=%3d  mean %9.1f us  total %10.1f us  %5.1f %%" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
    print("total %.1f us over %d launches" % (tot, len(data)))


if __name__ == "__main__":
    main(*sys.argv[1:])
