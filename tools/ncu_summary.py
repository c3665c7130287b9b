"""Reads .ncu-rep files (ncu -i, no GPU needed) and prints / stores the metrics the design discussion uses."""
import csv, json, subprocess, sys
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
 'smsp__thread_inst_executed_per_inst_executed.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'launch__grid_size','launch__block_size','sm__cycles_elapsed.avg','sm__cycles_active.avg',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']
def summary(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, v = rows[0], rows[1], rows[-1]
    d = {'kernel': v[hdr.index('Kernel Name')]}
    for k in KEYS:
        if k in hdr: d[k] = v[hdr.index(k)] + ' ' + units[hdr.index(k)]
    return d
if __name__ == '__main__':
    res = {p: summary(p) for p in sys.argv[1:]}
    names = list(res)
    for k in ['kernel'] + KEYS:
        print(f"{k[:84]:84s} " + ' | '.join(f"{res[n].get(k,'-')[:24]:>24s}" for n in names))
    json.dump(res, open('gpurun_out/ncu_summary_last.json','w'), indent=1)
