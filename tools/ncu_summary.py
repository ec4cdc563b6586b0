#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [out.tsv] — the handful of ncu counters this project argues from, per kernel."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__maximum_warps_per_active_cycle_pct', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fp64.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor']
out = []
for r in rows[2:]:
    out.append("# kernel\t" + r[h.index('Kernel Name')])
    for w in want:
        if w in h:
            out.append(f"{w}\t{units[h.index(w)]}\t{r[h.index(w)]}")
    st = []
    for i, name in enumerate(h):
        if name.startswith('smsp__pcsamp_warps_issue_stalled_') and not name.endswith('_not_issued') and r[i] not in ('', 'n/a'):
            st.append((float(r[i].replace(',', '')), name))
    tot = sum(v for v, _ in st) or 1
    for v, name in sorted(st, reverse=True)[:9]:
        out.append(f"{name}\tsamples\t{v:.0f}\t{v / tot:.3f}")
txt = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
print(txt)
