"""Key metrics of an `ncu --page raw --csv` export, one block per captured launch.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_summary.py raw.csv"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu.sum', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print('----')
        for w in WANT:
            if w in idx:
                print(f"{w:70s} {d[idx[w]][:90]} {units[idx[w]]}")
        for h in hdr:
            if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
                v = float(d[idx[h]])
                if v > 3:
                    print(f"   stall {h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):30s} {v:.1f} %")


if __name__ == '__main__':
    main(sys.argv[1])
