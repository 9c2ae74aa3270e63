#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the few metrics the roofline discussion needs.
Usage: python tools/ncu_summary.py file.ncu-rep [more metric substrings]"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum.per_second',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__shared_mem_per_block_static', 'launch__shared_mem_config_size',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'sm__ctas_launched.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_read.sum', 'smsp__inst_executed.sum']


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:] or ['issue_stalled']
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('%-90s %s %s' % (w, r[i], units[i]))
        for e in extra:
            for i, h in enumerate(hdr):
                if e in h and h not in WANT and (e != 'issue_stalled' or h.endswith('per_issue_active.ratio')):
                    print('%-90s %s %s' % (h, r[i], units[i]))
        print('-' * 60)


if __name__ == '__main__':
    main()
