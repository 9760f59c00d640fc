"""Summarise an `ncu --page raw --csv` dump: python tools/ncu_summary.py raw.csv [row]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max']
sel = [int(a) for a in sys.argv[2:]] or list(range(2, len(rows)))
for k in sel:
    r = rows[k]
    print('----- row', k)
    for w in want:
        for i, c in enumerate(h):
            if c == w:
                print("%-70s %s %s" % (w, r[i], rows[1][i]))
    for i, c in enumerate(h):
        if 'issue_stalled' in c and c.endswith('per_issue_active.ratio') and 'not_issued' not in c:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.15:
                print("   stall %-50s %s" % (c.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), r[i]))
