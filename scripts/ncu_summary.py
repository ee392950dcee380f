"""print the key metrics of an ncu --page raw --csv dump (one kernel per row)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_fp64', 'sm__pipe_fp64_cycles_active', 'smsp__inst_executed.sum ',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg ', 'lts__t_sectors_op_red.sum',
        'smsp__issue_active.avg.pct', 'smsp__warp_issue_stalled', 'local_load', 'local_store', 'lts__t_sector_hit_rate',
        'l1tex__t_sector_hit_rate', 'sm__inst_executed_pipe_lsu', 'smsp__cycles_active.avg ', 'fp64', 'dram__throughput',
        'lts__throughput', 'l1tex__throughput', 'sm__inst_executed.avg.per_cycle_elapsed', 'derived__smsp__sass_thread_inst_executed_op_dfma']
for r in rows[2:]:
    for i, h in enumerate(H):
        if any(k.strip() in h for k in keys):
            print("%-90s %-12s %s" % (h, rows[1][i], r[i]))
    print("-" * 40)
