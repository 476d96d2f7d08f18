"""Summarise an ncu report (raw + source pages exported as CSV) for profiles/."""
import csv
import sys
from collections import Counter

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, d = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
print('kernel:', d[hdr.index('Kernel Name')])
for k in keys:
    if k in hdr:
        print(f'  {k:70s} {d[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
print('stall samples (smsp__pcsamp_warps_issue_stalled_*):')
for i, k in enumerate(hdr):
    if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k:
        try:
            v = float(d[i])
        except ValueError:
            continue
        if v > 0:
            print(f'  {k.split("stalled_")[1]:24s} {int(v)}')
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]
data = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
ia, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
c, cs = Counter(), Counter()
for r in data:
    t = r[ia].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[iex] or 0)
    cs[op] += int(r[isamp] or 0)
nw = max(int(r[iex] or 0) for r in data if 'EXIT' not in r[ia]) if data else 1
tot = sum(c.values())
print(f'SASS: {len(data)} static instructions, {tot} executed warp-instructions')
for op, n in c.most_common(18):
    print(f'  {op:10s} executed {n:10d} ({100.0 * n / tot:5.1f}%)  samples {cs[op]}')
n = len(data)
B = 20
print('stall samples along the program (20 buckets of static instructions):')
il = h.index('stall_long_sb')
iw = h.index('stall_wait')
for b in range(B):
    seg = data[b * n // B:(b + 1) * n // B]
    s = sum(int(r[isamp] or 0) for r in seg)
    l = sum(int(r[il] or 0) for r in seg)
    w = sum(int(r[iw] or 0) for r in seg)
    print(f'  [{b * n // B:5d}..) samples {s:6d} long_sb {l:6d} wait {w:6d}')
