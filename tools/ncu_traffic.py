#!/usr/bin/env python
"""Summarise an ncu launch list (csv with dram__bytes_read.sum, dram__bytes_write.sum and
gpu__time_duration.sum per launch) into profiles/traffic.json: measured DRAM bytes per traversal
launch at the bench's own batch size, which bench.py reports as roofline.traffic.

usage: tools/ncu_traffic.py WORKLOAD launches.csv [source-note]
The csv comes from
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:k_split --csv --log-file launches.csv python bench.py --workload WORKLOAD --steps 1 --warmup 0 ...
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0,
        "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    workload, path = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    launches = {}
    for r in rows:
        key = r["ID"]
        d = launches.setdefault(key, {"kernel": r["Kernel Name"]})
        val = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        d[r["Metric Name"]] = val
    per_kernel = {}
    total_bytes = total_ms = 0.0
    n = 0
    for d in launches.values():
        name = d["kernel"].split("(")[0].replace("void ", "")
        if not name.startswith("k_split"):
            continue
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        ms = d.get("gpu__time_duration.sum", 0.0)
        k = per_kernel.setdefault(name, {"launches": 0, "bytes": 0.0, "ms": 0.0})
        k["launches"] += 1
        k["bytes"] += b
        k["ms"] += ms
        total_bytes += b
        total_ms += ms
        n += 1
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    table = {}
    if os.path.exists(out_path):
        with open(out_path) as f:
            table = json.load(f)
    table[workload] = {
        "bytes_per_launch": total_bytes / max(n, 1), "launches": n, "ms_under_ncu": total_ms,
        "metric": "dram__bytes_read.sum + dram__bytes_write.sum, averaged over the traversal launches captured",
        "per_kernel": {k: {"launches": v["launches"], "bytes_per_launch": v["bytes"] / v["launches"],
                           "ms_per_launch_under_ncu": v["ms"] / v["launches"]} for k, v in sorted(per_kernel.items())},
        "source": os.path.relpath(path, ROOT), "note": note,
    }
    with open(out_path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(json.dumps(table[workload], indent=1))


if __name__ == "__main__":
    main()
