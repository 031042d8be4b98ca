"""Summarise an .ncu-rep (raw page) and a launch list CSV into text for profiles/."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
launch = sys.argv[2] if len(sys.argv) > 2 else None
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum', 'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'smsp__inst_executed_op_global_atom.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__thread_inst_executed_per_inst_executed.ratio']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
idx = {n: i for i, n in enumerate(h)}
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0].replace('void <unnamed>::', '').replace('<unnamed>::', '')
    print(f"== {name}")
    for w in WANT:
        if w in idx and r[idx[w]] != '':
            print(f"   {w:78s} {r[idx[w]]:>18s} {units[idx[w]]}")
if launch:
    rows = list(csv.reader(open(launch)))
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hh = rows[hdr]; ki = hh.index('Kernel Name'); vi = hh.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split('(')[0].replace('void <unnamed>::', '').replace('<unnamed>::', ''), []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print("== launch list (gpu__time_duration, cold-cache serialised: compare shares)")
    for k, v in agg.items():
        print(f"   {k[:60]:60s} n={len(v):3d} mean_us={sum(v)/len(v)/1e3:9.1f} share={sum(v)/tot*100:5.1f}%")
