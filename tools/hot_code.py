#!/usr/bin/env python3
"""How much code holds the hot path of a kernel?  Weights = warp instructions per source line from an Nsight Compute report (source page),
sizes = SASS instructions per source line of the CURRENT build (nvdisasm -g on the in-tree library): lines are taken hottest first until they
cover 50 / 90 / 99 % of the profiled instructions, and the bytes of code behind them are summed (16 B per instruction).  The SM's instruction
caches are ~6 KB (L0, per scheduler) and 32 KB (L1.5): a hot set far above that is fetched from the GPC-level cache over and over.
Usage: tools/hot_code.py report.ncu-rep kernel_name [library.so]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'opus_b200', 'libopus_b200.so')
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
dyn = collections.Counter(); cur = None; h = None
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] in ('File Path', 'File Name'):
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        h = r; continue
    if not h or not r or r[0] in ('', 'Function Name'):
        continue
    try:
        dyn[(cur, int(r[0]))] += int(r[7])   # the source-level row of the mixed view: column 7 = warp instructions of the line
    except (ValueError, IndexError):
        pass
with tempfile.TemporaryDirectory() as d:
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
    sass = subprocess.run(['nvdisasm', '-g', os.path.join(d, cub)], capture_output=True, text=True).stdout
stat = collections.Counter(); inside = False; line = None; total = 0
for l in sass.splitlines():
    if l.startswith('.text.'):
        inside = l.strip() == '.text.%s:' % kern; continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l):
        stat[line] += 1; total += 1
td = sum(dyn.values()) or 1
print('%s: %d SASS instructions = %.1f KB in this build; profile: %d warp instructions on %d source lines' % (kern, total, total * 16 / 1024, td, len(dyn)))
acc = 0; size = 0; marks = [0.5, 0.9, 0.99]
for k, v in dyn.most_common():
    acc += v; size += stat.get(k, 0)
    while marks and acc >= marks[0] * td:
        print('  %2.0f %% of the profiled instructions come from lines holding %6.1f KB of code' % (100 * marks[0], size * 16 / 1024)); marks.pop(0)
    if not marks:
        break
hot = [(stat.get(k, 0), v, k) for k, v in dyn.most_common() if stat.get(k, 0)]
acc = 0; sel = []
for st_, v, k in sorted(hot, key=lambda t: -t[1]):
    acc += v; sel.append((st_, v, k))
    if acc >= 0.9 * td:
        break
reg = collections.Counter(); regd = collections.Counter()
for st_, v, k in sel:
    reg[(k[0], k[1] // 25 * 25)] += st_; regd[(k[0], k[1] // 25 * 25)] += v
print('  code behind the 90 % set, by 25-line region (KB, % of profiled instructions):')
for k, v in reg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 25):
    print('    %5.1f KB  %4.1f %%  %s:%d' % (v * 16 / 1024, 100.0 * regd[k] / td, k[0], k[1]))
