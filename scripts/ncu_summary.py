"""Print the key per-launch metrics of an .ncu-rep (raw page) -- used to write profiles/*.txt."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed' ,
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__icc_request_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for v in rows[2:]:
    d = dict(zip(h, v))
    print('  %-70s %s' % ('Kernel Name', d.get('Kernel Name', '')[:110]))
    for k in KEYS:
        for kk in h:
            if kk == k or kk.endswith('.' + k): print('  %-70s %s %s' % (k, d[kk], rows[1][h.index(kk)])); break
    st = sorted(((float(d[k]), k) for k in h if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k), reverse=True)
    print('  stall cycles per issued instruction:', ', '.join('%s %.2f' % (k.split('issue_stalled_')[1].split('_per_issue')[0], x) for x, k in st[:8]))
