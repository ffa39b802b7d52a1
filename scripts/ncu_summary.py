#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: the metrics DESIGN.md/profiles cite."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index('Kernel Name')][:80], "id", r[0])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-70s %s %s" % (w, r[i], units[i]))
        st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
              if h.startswith('smsp__average_warp') and 'per_issue_active' in h and r[i] not in ('', 'n/a')]
        st = st or [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
                    if h.startswith('smsp__warp_issue_stalled') and h.endswith('per_warp_active.pct') and r[i] not in ('', 'n/a')]
        for v, h in sorted(st, reverse=True)[:8]:
            print("  stall %-64s %.2f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', ''), v))


if __name__ == "__main__":
    main(sys.argv[1])
