"""Summarise an `ncu --page raw --csv` export: one block per launch with the metrics the roofline
discussion needs + the top stall reasons.   python profiles/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for d in data:
    print('=====', d[idx['Kernel Name']][:110])
    for w in want:
        if w in idx:
            print('  %-62s %18s %s' % (w, d[idx[w]][:18], units[idx[w]]))
    st = []
    for h, i in idx.items():
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
            try:
                st.append((float(d[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError:
                pass
    print('  stalls per issue:', ', '.join('%s %.2f' % (n, v) for v, n in sorted(st, reverse=True)[:6]))
