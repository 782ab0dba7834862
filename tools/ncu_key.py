"""Key metrics of an `ncu --page raw --csv` export.  usage: ncu_key.py <raw.csv> [points]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
pts = float(sys.argv[2]) if len(sys.argv) > 2 else 148 * 347200
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
for w in want:
    if w in d: print('%-75s %s %s' % (w, d[w][0], d[w][1]))
if 'smsp__inst_executed.sum' in d:
    print('warp instructions per 64 points: %.1f' % (float(d['smsp__inst_executed.sum'][0]) / pts * 64))
for h in hdr:
    if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_warp_active.pct'):
        v = float(d[h][0])
        if v > 2: print('  stall %-60s %.1f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
