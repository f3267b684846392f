"""Development aid: compact digest of an `ncu --set full` capture for profiles/.
    ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu -i X.ncu-rep --page source --csv > src.csv
    python tools/ncu_digest.py raw.csv [src.csv] [units] > profiles/rNN_<kernel>.md
Prints, per captured kernel launch, the counters the design argues from (duration, FP64-pipe and issue activity,
shared-memory wavefronts, DRAM bytes, registers, top stall reasons) and, with a source page, the SASS opcode mix."""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__warps_active.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("## %s" % r[hdr.index("Kernel Name")])
        print()
        print("| metric | value | unit |")
        print("|---|---|---|")
        for k in KEYS:
            if k in hdr:
                print("| %s | %s | %s |" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        st = []
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                st.append((float(r[i] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        print()
        print("stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
        print()
    if len(sys.argv) > 2:
        src = list(csv.reader(open(sys.argv[2])))
        units_n = float(sys.argv[3]) if len(sys.argv) > 3 else None
        h = src[1]
        isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        mix = collections.Counter(); smp = collections.Counter(); tot = 0
        for r in src[2:]:
            if len(r) <= iex:
                continue
            toks = r[isrc].split()
            op = toks[0] if not toks[0].startswith("@") else toks[1]
            op = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDG", "STG", "RED")) else ".".join(op.split(".")[:2])
            n = int(r[iex] or 0)
            mix[op] += n; tot += n; smp[op] += int(r[ismp] or 0)
        ts = max(sum(smp.values()), 1)
        print("### SASS opcode mix (warp instructions executed%s)" % (", per warp-level unit" if units_n else ""))
        print()
        print("| opcode | executed | share | per unit | stall samples |")
        print("|---|---|---|---|---|")
        for op, n in mix.most_common(16):
            print("| %s | %.4g | %.1f %% | %s | %.1f %% |" % (op, n, 100.0 * n / tot, ("%.2f" % (n / units_n)) if units_n else "", 100.0 * smp[op] / ts))
        print("| total | %.4g | | %s | |" % (tot, ("%.1f" % (tot / units_n)) if units_n else ""))


if __name__ == "__main__":
    main()
