#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of counters DESIGN.md cites."""
import csv, subprocess, sys
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_fmaheavy.sum',
        'sm__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_adu.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_global_st.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']
STALL = 'smsp__average_warps_issue_stalled_'
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    w = csv.writer(sys.stdout)
    w.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(data))])
    for name in WANT + sorted(h for h in hdr if h.startswith(STALL) and h.endswith('_per_issue_active.ratio')):
        if name in hdr:
            i = hdr.index(name)
            w.writerow([name, units[i]] + [r[i] for r in data])
if __name__ == '__main__':
    main(sys.argv[1])
