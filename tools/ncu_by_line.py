#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line, using the
line table nvdisasm prints for the same cubin (build with -lineinfo).

    cuobjdump -xelf all librayito_b200.so ; nvdisasm -g -c rt_core.sm_100a.cubin > dis.txt
    ncu -i prof.ncu-rep --page source --csv --kernel-id :::N > sass.csv
    python tools/ncu_by_line.py dis.txt sass.csv '<mangled kernel name>' [top] [section]

section: when the csv holds several launches (one "Kernel Name" block per launch and view), use only the
block with this index (0-based); default: all blocks (they must then be launches of the same kernel).
"""
import csv
import collections
import re
import sys


def line_table(dis_path, mangled):
    table, cur, active = {}, None, False
    for raw in open(dis_path):
        if raw.startswith(".text."):
            active = raw.strip() == ".text.%s:" % mangled
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', raw)
        if m:
            # keep only the innermost location (first one printed before the instruction)
            if cur is None or "inlined at" not in raw:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", raw)
        if m and cur is not None:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    dis_path, csv_path, mangled = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(dis_path, mangled)
    rows = list(csv.reader(open(csv_path)))
    if len(sys.argv) > 5:
        starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
        k = int(sys.argv[5])
        rows = rows[starts[k]:starts[k + 1]]
        print("# section %d: %s" % (k, rows[0][1][:100]))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {name: hdr.index(name) for name in ("Address", "Source", "# Samples", "Instructions Executed",
                                              "Thread Instructions Executed", "stall_long_sb", "stall_wait",
                                              "stall_short_sb", "stall_math", "stall_branch_resolving", "stall_no_inst",
                                              "stall_not_selected", "stall_lg", "stall_dispatch")}
    agg = collections.defaultdict(lambda: collections.Counter())
    base = None
    for r in rows[hdr_i + 1:]:
        if r and r[0] == "Kernel Name":
            base = None             # next launch of the same kernel: addresses restart
            continue
        if len(r) < len(hdr) - 1 or r[0] == "Address":
            continue
        addr = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
        if base is None:
            base = addr
        loc = table.get(addr - base, ("?", 0))
        a = agg[loc]
        for k in col:
            if k in ("Address", "Source"):
                continue
            try:
                a[k] += float(r[col[k]].replace(",", "") or 0)
            except ValueError:
                pass
        a["sass"] += 1
    tot = collections.Counter()
    for a in agg.values():
        tot.update(a)
    print("total: samples %d, warp instr %.3g, thread instr %.3g, avg active lanes %.2f" % (
        tot["# Samples"], tot["Instructions Executed"], tot["Thread Instructions Executed"],
        tot["Thread Instructions Executed"] / max(tot["Instructions Executed"], 1)))
    stall_names = [k for k in col if k.startswith("stall_")]
    print("stalls: " + ", ".join("%s %.1f%%" % (k[6:], 100 * tot[k] / max(tot["# Samples"], 1)) for k in stall_names))
    print("%-28s %6s %8s %8s %6s  top stalls" % ("file:line", "sass", "instr%", "samp%", "lanes"))
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        stalls = sorted(((a[k], k[6:]) for k in stall_names), reverse=True)[:3]
        print("%-28s %6d %7.2f%% %7.2f%% %6.1f  %s" % (
            "%s:%d" % loc, a["sass"], 100 * a["Instructions Executed"] / max(tot["Instructions Executed"], 1),
            100 * a["# Samples"] / max(tot["# Samples"], 1),
            a["Thread Instructions Executed"] / max(a["Instructions Executed"], 1),
            " ".join("%s=%.0f" % (n, v) for v, n in stalls if v)))


if __name__ == "__main__":
    main()
