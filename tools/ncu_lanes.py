#!/usr/bin/env python3
"""Per source line of an Nsight Compute report: warp instructions, average active lanes, and the issue slots lost to idle lanes
(instructions x (32 - lanes) / 32) — where SIMT divergence costs.  Also sums per function-ish file region.
Usage: tools/ncu_lanes.py report.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, h = None, None
lines = []
for r in rows:
    if len(r) == 2 and r[0] in ('File Path', 'File Name'):
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        h = r; continue
    if not h or not r or r[0] in ('', 'Function Name'):
        continue
    try:
        ln = int(r[0]); inst = int(r[7]); tinst = int(r[8]); samp = int(r[6])
    except (ValueError, IndexError):
        continue
    if inst:
        lines.append((cur, ln, inst, tinst, samp, r[1].strip()[:90]))
ti = sum(l[2] for l in lines); tt = sum(l[3] for l in lines); ts = sum(l[4] for l in lines) or 1
print('warp instructions %d, thread instructions %d -> %.2f lanes on average' % (ti, tt, tt / ti))
waste = sorted(lines, key=lambda l: -(l[2] * 32 - l[3]))
tw = sum(l[2] * 32 - l[3] for l in lines)
print('== lines by idle-lane issue slots (%% of all idle slots, %% of instructions, lanes, %% samples)')
for l in waste[:top]:
    print('  %5.1f%% %5.1f%% %5.1f %5.1f%%  %s:%d  %s' % (100.0 * (l[2] * 32 - l[3]) / tw, 100.0 * l[2] / ti, l[3] / l[2], 100.0 * l[4] / ts, l[0], l[1], l[5]))
byfile = collections.defaultdict(lambda: [0, 0, 0])
for l in lines:
    k = (l[0], l[1] // 50 * 50)
    byfile[k][0] += l[2]; byfile[k][1] += l[3]; byfile[k][2] += l[4]
print('== 50-line regions by instructions')
for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0])[:top]:
    print('  %5.1f%% inst  %5.1f lanes  %5.1f%% samples  %s:%d-%d' % (100.0 * v[0] / ti, v[1] / v[0], 100.0 * v[2] / ts, k[0], k[1], k[1] + 49))
