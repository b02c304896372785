#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + per-opcode counts/stall samples from the source page). CPU-only."""
import collections, csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']


def main(path, kfilter=None, top=14):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        if kfilter and kfilter not in name:
            continue
        print('==', name[:110])
        for w in WANT:
            if w in idx:
                print(f'   {w:66s} {r[idx[w]]:>16s} {units[idx[w]]}')
        stall = [(h, r[idx[h]]) for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h]
        st = sorted(((float(v), h) for h, v in stall if v not in ('', 'n/a')), reverse=True)[:8]
        for v, h in st:
            print(f'   stall {h.split("issue_stalled_")[1][:40]:42s} {v:8.2f} warp-cycles/instr')
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    kern, hd = None, None
    cnt = collections.defaultdict(collections.Counter); smp = collections.defaultdict(collections.Counter)
    seen_addr = collections.defaultdict(set)
    nk = collections.Counter()
    for row in csv.reader(io.StringIO(src)):
        if not row:
            continue
        if row[0] == 'Kernel Name':
            kern = row[1]; nk[kern] += 1; hd = None; continue
        if row[0] == 'Address':
            hd = row; continue
        if hd is None or (kfilter and kfilter not in kern) or nk[kern] > 1:
            continue
        if row[0] in seen_addr[kern]:
            continue
        seen_addr[kern].add(row[0])
        sass = row[1].strip().split()
        op = (sass[1] if sass[0].startswith('@') else sass[0]).split('.')[0]
        cnt[kern][op] += int(row[hd.index('Instructions Executed')])
        smp[kern][op] += int(row[hd.index('# Samples')])
    for k, c in cnt.items():
        tot = sum(c.values()); ts = sum(smp[k].values()) or 1
        print('-- opcode mix', k[:90], 'warp-instr', tot)
        for op, n in c.most_common(top):
            print(f'   {op:10s} {n:11d} {100.0*n/tot:6.2f}%   stall samples {100.0*smp[k][op]/ts:6.2f}%')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
