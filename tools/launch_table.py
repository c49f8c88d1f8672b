#!/usr/bin/env python
"""Print the launches of one wavefront batch from an ncu launch-list CSV
(gpu__time_duration + lanes + instructions + issue/warps active)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 7
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mn, mv, idc = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > mv:
        per.setdefault(r[idc], {'name': r[kn].split('(')[0][:44]})[r[mn]] = float(r[mv].replace(',', ''))
lst = list(per.values())
ray_idx = [i for i, k in enumerate(lst) if k['name'] == 'k_raygen']
batch = min(batch, len(ray_idx) - 1)       # the last batch runs to the end of the list
start = ray_idx[batch] - 1
end = ray_idx[batch + 1] - 1 if batch + 1 < len(ray_idx) else len(lst)
tot = collections.Counter()
for k in lst[start:end]:
    t = k['gpu__time_duration.sum'] / 1e3
    tot[k['name']] += t
    if t > 40:
        print("%-46s %8.1f us lanes %5.1f inst %.3g issue %4.1f%% warps %4.1f%%" % (
            k['name'], t, k['smsp__thread_inst_executed_per_inst_executed.ratio'], k['smsp__inst_executed.sum'],
            k['smsp__issue_active.avg.pct_of_peak_sustained_active'], k['sm__warps_active.avg.pct_of_peak_sustained_active']))
print('--- batch total %.1f us' % sum(tot.values()))
for n, t in tot.most_common():
    print("%-46s %8.1f us" % (n, t))
