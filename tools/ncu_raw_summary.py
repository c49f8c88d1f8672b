#!/usr/bin/env python
"""Compact per-launch summary of an `ncu --set full ... --page raw --csv` export: the handful of counters the
profiles/README.md tables quote (time, lanes per instruction, issue slots busy, DRAM bytes, hit rates, warps
resident, registers, the three largest stall reasons in cycles per issued instruction).

usage: tools/ncu_raw_summary.py raw.csv > summary.csv"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes_per_inst"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(hdr)}
    stalls = [n for n in hdr if n.startswith(STALL_PREFIX) and n.endswith(STALL_SUFFIX)]
    out = csv.writer(sys.stdout)
    present = [(n, short) for n, short in COLS if n in col]
    out.writerow(["kernel"] + ["%s [%s]" % (short, units[col[n]]) if units[col[n]] else short for n, short in present] +
                 ["top stalls (cycles per issued instruction)"])
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(DScene")[0].split("(RenderCtx")[0].replace("void ", "")
        vals = []
        for n, _short in present:
            v = r[col[n]].replace(",", "")
            try:
                vals.append("%.4g" % float(v))
            except ValueError:
                vals.append(v)
        st = []
        for n in stalls:
            try:
                st.append((float(r[col[n]].replace(",", "")), n[len(STALL_PREFIX):-len(STALL_SUFFIX)]))
            except ValueError:
                pass
        st.sort(reverse=True)
        out.writerow([name] + vals + ["; ".join("%s %.2f" % (n, v) for v, n in st[:3])])


if __name__ == "__main__":
    main()
