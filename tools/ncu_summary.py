"""Summarise an ncu --set full report: key raw metrics + top stall instructions. usage: ncu_summary.py rep [topN] [kernel-regex]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ksel = ["-k", "regex:" + sys.argv[3], "-c", "1"] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + ksel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'lts__t_sectors_srcunit_tex.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.per_cycle_active']
for i, name in enumerate(h):
    if name in want or ('issue_stalled' in name and 'per_issue_active' in name and float(v[i] or 0) > 0.15):
        print(f'{name:90s} {u[i]:15s} {v[i]}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + ksel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
si = h.index('# Samples'); sc = h.index('Source'); ie = h.index('Instructions Executed')
cols = {k: h.index(k) for k in ['stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving', 'stall_math', 'stall_mio']}
tot = sum(int(r[si]) for r in data)
print('total samples', tot, 'n instr', len(data))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:topn]
for i in sorted(idx):
    r = data[i]
    print(f'{i:5d} {int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}% exec={r[ie]:>9s} ' + ' '.join(f'{k[6:9]}={r[c]:>5s}' for k, c in cols.items()) + f' {r[sc][:80]}')
