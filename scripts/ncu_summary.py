"""Summarise an .ncu-rep: key metrics, stall reasons, SASS mix, hottest source lines."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            print(f"{w} = {r[hdr.index(w)]} {units[hdr.index(w)]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
            try: st.append((float(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError: pass
    print("stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(rows):
    if 'Source' in r and 'Instructions Executed' in r:
        h = r; start = i + 1; break
else:
    sys.exit(0)
si, ei, sa = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
mix, tot = Counter(), 0
lines = []
for r in rows[start:]:
    if len(r) <= max(ei, sa): continue
    try: n = int(r[ei]); smp = int(r[sa])
    except ValueError: continue
    toks = r[si].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    mix[op.split('.')[0]] += n; tot += n
    lines.append((smp, n, r[si][:90]))
print("SASS mix (warp instructions):", tot)
print(", ".join(f"{k} {100*v/tot:.1f}%" for k, v in mix.most_common(16)))
print("hottest SASS by stall samples:")
ts = sum(l[0] for l in lines) or 1
for smp, n, s in sorted(lines, reverse=True)[:14]:
    print(f"  {100*smp/ts:5.1f}%  exec={n:10d}  {s}")
