"""Print the metrics we track from an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'launch__grid_size','launch__block_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio','smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h:i for i,h in enumerate(hdr)}
    print('==', rep)
    print('kernel', [r[idx['Kernel Name']][:60] for r in data])
    for w in WANT:
        if w in idx:
            print(f"  {w} [{units[idx[w]]}]", [r[idx[w]] for r in data])
