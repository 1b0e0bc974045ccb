#!/usr/bin/env python3
"""Summarise an Nsight Compute report (no GPU needed): key raw metrics, stall-reason mix and the
hottest source lines.  Usage: tools/ncu_summary.py report.ncu-rep [n_frames] [top]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'sm__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gcc__cache_requests_type_instruction.sum.pct', 'sm__icc_request_hit_rate', 'l1tex__t_sector_hit_rate.pct',
        'mem_local_op_ld.sum', 'mem_local_op_st.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nfr = float(sys.argv[2]) if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, vals = raw[0], raw[-1]
    print('== raw metrics')
    for h, v in zip(hdr, vals):
        if any(k in h for k in KEYS) and 'per_second' not in h and '.max' not in h and '.min' not in h:
            print('  %-90s %s' % (h, v))
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv']))))
    cur, h = None, None
    stalls = collections.Counter()
    lines = []
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            h = r
            continue
        if not h or not r:
            continue
        if r[0] == '' and len(r) > 10 and r[2].startswith('0x'):
            for i, name in enumerate(h):
                if name.startswith('stall_') and 'Not Issued' not in name:
                    try:
                        stalls[name] += int(r[i])
                    except ValueError:
                        pass
        elif r[0] not in ('', 'Function Name'):
            try:
                lines.append((int(r[6]), int(r[7]), cur, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    tot = sum(stalls.values()) or 1
    print('== stall reasons (all samples)')
    for k, v in stalls.most_common(9):
        print('  %-26s %5.1f%%' % (k, 100.0 * v / tot))
    ts = sum(a[0] for a in lines) or 1
    ti = sum(a[1] for a in lines) or 1
    print('== warp instructions executed: %d%s' % (ti, ('  = %.0f per frame' % (ti / nfr)) if nfr else ''))
    print('== hottest source lines (%% samples, %% instructions)')
    for a in sorted(lines, reverse=True)[:top]:
        print('  %5.1f%% %5.1f%%  %s:%d  %s' % (100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2], a[3], a[4]))


if __name__ == '__main__':
    main()
