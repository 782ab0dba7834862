#!/bin/bash
# Read an .ncu-rep here (no GPU): key metrics + per-segment instruction counts per 64-point warp tile.
# usage: tools/ncu_summary.sh gpurun_out/<tag>.ncu-rep <tiles>
rep=$1; tiles=${2:-803640}; base=${rep%.ncu-rep}
ncu -i $rep --page raw --csv > ${base}_raw.csv 2>/dev/null
ncu -i $rep --page source --csv > ${base}_src.csv 2>/dev/null
python - "$base" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1] + '_raw.csv')))
hdr, vals = rows[0], rows[-1]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread']
for i, h in enumerate(hdr):
    if h in want: print(f"{h:90s} {vals[i]}")
PY
python tools/ncu_segments.py ${base}_src.csv $tiles ${3:-4}
