"""Reads .ncu-rep files (ncu -i, no GPU needed) and prints / stores the metrics the design discussion uses.
Usage: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]   (every profiled launch of every report becomes one column)"""
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
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum']
def summaries(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for v in rows[2:]:
        d = {'kernel': v[hdr.index('Kernel Name')][:60]}
        for k in KEYS:
            if k in hdr: d[k] = v[hdr.index(k)] + ' ' + units[hdr.index(k)]
        res.append(d)
    return res
if __name__ == '__main__':
    res = {}
    for p in sys.argv[1:]:
        for i, d in enumerate(summaries(p)):
            res['%s#%d' % (p.split('/')[-1].replace('.ncu-rep', ''), i)] = d
    names = list(res)
    print(f"{'':84s} " + ' | '.join(f"{n[-24:]:>24s}" for n in names))
    for k in ['kernel'] + KEYS:
        print(f"{k[:84]:84s} " + ' | '.join(f"{res[n].get(k,'-')[-24:]:>24s}" for n in names))
    json.dump(res, open('gpurun_out/ncu_summary_last.json','w'), indent=1)
