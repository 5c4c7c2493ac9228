#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, dynamic instruction mix and top stall sites.
usage: python profiles/summarise.py gpurun_out/prof.ncu-rep [kernel-regex] > profiles/summary.txt"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_srcunit_tex.sum', 'lts__t_bytes.sum',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    print('=== kernel:', r[hdr.index('Kernel Name')][:110])
    for w in WANT:
        if w in hdr:
            print(f'  {w:75s} {r[hdr.index(w)]} {rows[1][hdr.index(w)]}')
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(r[i] or 0) > 0.05:
            print(f'  {h:75s} {r[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
iA, iS, iN, iE = (hdr.index(k) for k in ('Address', 'Source', '# Samples', 'Instructions Executed'))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base = int(data[0][iA], 16)
mix = collections.Counter()
for r in data:
    t = re.sub(r'^@!?U?P\d\s+', '', r[iS].strip())
    mix[t.split()[0].split('.')[0]] += int(r[iE])
loop = max(int(r[iE]) for r in data if 'FFMA' in r[iS])
tot = sum(mix.values())
print(f'\n=== dynamic instruction mix (first captured launch): total {tot}, per main-loop trip {tot / loop:.1f}')
print('  ' + '  '.join(f'{k}:{v / loop:.1f}' for k, v in mix.most_common(24)))
totN = sum(int(r[iN]) for r in data)
print(f'\n=== top stall sites ({totN} samples)')
for r in sorted(data, key=lambda r: -int(r[iN]))[:22]:
    st = {hdr[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"  {int(r[iA], 16) - base:5x} {100 * int(r[iN]) / totN:5.1f}%  {r[iS].strip()[:58]:58s} {st}")
