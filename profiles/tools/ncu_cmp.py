"""side-by-side of the raw metrics that matter for the coder: python profiles/tools/ncu_cmp.py a.ncu-rep b.ncu-rep ..."""
import csv, subprocess, sys, io
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__block_size','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
'l1tex__lsu_writeback_active_mem_lgds.sum.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed',
'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg','SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg','SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg','sm__cycles_elapsed.max']
cols=[]
for rep in sys.argv[1:]:
    raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(io.StringIO(raw))); cols.append(dict(zip(rows[0],rows[2])))
for k in keys:
    print('%-95s'%k[:95], *['%16s'%c.get(k,'-')[:16] for c in cols])
