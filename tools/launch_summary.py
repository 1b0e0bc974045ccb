#!/usr/bin/env python3
"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum, sm__inst_executed.sum, smsp__thread_inst_executed_per_inst_executed.ratio):
per kernel the launches, the last launch's time, warp instructions and average active lanes.  Usage: tools/launch_summary.py list.csv [filter]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
flt = sys.argv[2] if len(sys.argv) > 2 else ''
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        h, start = r, i + 1
        break
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) < len(h):
        continue
    d = dict(zip(h, r))
    agg.setdefault(d['Kernel Name'], collections.defaultdict(list))[d['Metric Name']].append(float(d['Metric Value'].replace(',', '')))
tot = 0.0
for k, m in agg.items():
    if flt not in k:
        continue
    t = m['gpu__time_duration.sum']
    ins = m.get('sm__inst_executed.sum', [0])[-1]
    ln = m.get('smsp__thread_inst_executed_per_inst_executed.ratio', [0])[-1]
    tot += t[-1]
    print('%-34s launches %3d  last %8.3f ms  warp-inst %8.1f M  lanes %4.1f' % (k[:34], len(t), t[-1] / 1e6, ins / 1e6, ln))
print('sum of last launches: %.3f ms' % (tot / 1e6))
