#!/usr/bin/env python
""" Condenses `ncu -i X.ncu-rep --page raw --csv` exports into the few counters DESIGN.md / bench.py quote (per launch): duration, DRAM bytes read +
written (the roofline `traffic`), tensor-pipe activity, issue activity, registers, grid. Usage: ncu_summary.py raw.csv [raw2.csv ...] """
import csv
import sys

WANT = [('gpu__time_duration.sum', 'us'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs'),
        ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_%'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_%'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_%'),
        ('smsp__inst_executed.sum', 'warp_insts'), ('launch__waves_per_multiprocessor', 'waves')]

for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print(f'# {path}')
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0]
        parts = []
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                parts.append(f'{short}={r[i]}{(" " + units[i]) if short in ("dram_rd", "dram_wr") else ""}')
        print(f'{name}: ' + ', '.join(parts))
