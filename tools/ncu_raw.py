"""Print the interesting raw metrics (and the warp-stall breakdown) of every kernel in an .ncu-rep file."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
for r in rows[2:]:
    for k in keys:
        if k in h:
            print("%-80s %s %s" % (k, r[h.index(k)], u[h.index(k)]))
    st = []
    for i, k in enumerate(h):
        if 'average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio') or \
           ('average_warp_latency_issue_stalled' in k and k.endswith('.ratio')):
            try:
                st.append((float(r[i]), k))
            except ValueError:
                pass
    for v, k in sorted(st, reverse=True)[:10]:
        print("    stall %-70s %.2f" % (k.replace('smsp__average_', ''), v))
    print()
