#!/usr/bin/env python
"""Development aid: stall samples of a warp-specialised kernel split at its USETMAXREG instructions (roles), from
`ncu -i X.ncu-rep --page source --csv > src.csv`.   python tools/ncu_roles.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
cuts = [i for i, r in enumerate(data) if "USETMAXREG" in r[ix["Source"]]]
bounds = [0] + cuts + [len(data)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for k in range(len(bounds) - 1):
    lo, hi = bounds[k], bounds[k + 1]
    s = sum(f(r, "# Samples") for r in data[lo:hi]); ie = sum(f(r, "Instructions Executed") for r in data[lo:hi])
    print("region %d [%d, %d): samples %.0f (%.1f %%), warp instructions executed %.3e" % (k, lo, hi, s, 100 * s / tot, ie))
    print("   ", {h[6:]: round(sum(f(r, h) for r in data[lo:hi]) / tot * 100, 1) for h in stalls if sum(f(r, h) for r in data[lo:hi]) / tot > 0.005})
top = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:ntop]
for i in sorted(top):
    r = data[i]
    st = {h[6:]: int(f(r, h)) for h in stalls if f(r, h) > 0.15 * f(r, "# Samples")}
    print(i, r[ix["Source"]][:72].ljust(72), int(f(r, "# Samples")), int(f(r, "Instructions Executed")), st)
