"""Development aid: opcode mix of a kernel from an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv).
python tools/sass_mix.py file.csv [units]   -- `units` = warp-level work items, to print instructions per unit."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix = collections.Counter(); smp = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iex:
        continue
    toks = r[isrc].split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDG", "STG", "RED")) else ".".join(op.split(".")[:2])
    n = int(r[iex] or 0)
    mix[op] += n; tot += n; smp[op] += int(r[ismp] or 0)
ts = sum(smp.values())
print("total warp instructions %.4g" % tot + ("  per unit %.1f" % (tot / units) if units else ""))
for op, n in mix.most_common(28):
    print("%-12s %12.4g  %5.1f%%  %s  samples %4.1f%%" % (op, n, 100.0 * n / tot, ("%6.2f/unit" % (n / units)) if units else "", 100.0 * smp[op] / max(ts, 1)))
