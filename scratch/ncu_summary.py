import csv, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H, U = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second', 'sm__pipe_tensor_cycles_active_realtime.avg.pct',
        'sm__pipe_tensor_subpipe_hmma', 'gpu__dram_throughput.avg.pct', 'lts__throughput.avg.pct', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread ',
        'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'sm__cycles_elapsed.avg ', 'sm__cycles_elapsed.avg.per_second', 'sm__throughput.avg.pct', 'launch__occupancy_limit',
        'smsp__inst_executed.sum ', 'sm__warps_active.avg.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg ', 'sm__pipe_fp64_cycles_active.avg.pct', 'sm__inst_executed_pipe_fp64',
        'smsp__pcsamp_warps_issue_stalled', 'local_', 'smsp__warp_issue_stalled_long_scoreboard', 'sm__pipe_fma_cycles_active.avg.pct','smsp__issue_active.avg.pct', 'derived__smsp__sass_thread_inst_executed_op_dfma','l1tex__t_bytes_pipe_lsu_mem_local']
print(title)
for i, h in enumerate(H):
    if any(h.startswith(k) for k in keys):
        print(f"{h} [{U[i]}]: " + ", ".join(r[i] for r in rows[2:]))
