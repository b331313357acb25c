"""Key metrics per launch out of an `ncu --page raw --csv` dump.  usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python ncu_raw_summary.py raw.csv [every_nth]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
step = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct']
for n, r in enumerate(rows[2:]):
    if n % step != step - 1:
        continue
    print('--- launch', n, r[idx['Kernel Name']][:70], r[idx['Grid Size']])
    for k in keys:
        if k in idx:
            print(f"  {k} [{units[idx[k]]}] = {r[idx[k]]}")
