#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): key raw metrics of the first captured launch + stall samples per
SASS region.   usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [chunk]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 400
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__cycles_elapsed.max']
for d in rows[2:]:
    print('==', d[hdr.index('Kernel Name')][:90])
    for w in WANT:
        if w in hdr: print(f'  {w:78s} {d[hdr.index(w)]:>16s} {units[hdr.index(w)]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr = rows[1]
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] == 'Address': break   # first captured launch only
        data.append(r)
    isamp, iex, isrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
    names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[isamp]) for r in data) or 1
    print('SASS instructions', len(data), 'stall samples', tot)
    for c in range(0, len(data), chunk):
        ch = data[c:c + chunk]
        s = sum(int(r[isamp]) for r in ch)
        if not s: continue
        br = sorted(((sum(int(r[hdr.index(n)]) for r in ch), n) for n in names), reverse=True)[:3]
        print(f'  instr {c:6d}: {100 * s / tot:5.1f}% of samples, executed {sum(int(r[iex]) for r in ch):11d}  top', [(n, v) for v, n in br])
    print('  hottest instructions:')
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:12]:
        print(f'   {data.index(r):6d} {r[isamp]:>6s}  {r[isrc].strip()[:100]}')
