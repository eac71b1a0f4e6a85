"""key metrics of every kernel in an ncu report: python tools/ncu_key.py X.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg',
        'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr:
            v = r[hdr.index(w)]
            try:
                if float(v) == 0: continue
            except ValueError: pass
            print('   %-85s %s' % (w.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', ''), v))
