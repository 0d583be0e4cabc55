#!/usr/bin/env python
""" CUDA-source-line digest of an `ncu --set full --import-source on` capture (kernels compiled with -lineinfo):

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME --launch-count 1 > src.csv
    python tools/ncu_srclines.py src.csv [top_n]

Prints, per source line, the warp instructions executed and the warp-stall samples attributed to it (largest first), and the totals: which lines
of the kernel the instruction count and the stalls come from. """
import csv
import sys


def to_int(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path, errors='replace')))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No' and len(r) > 4)
    hdr = rows[hdr_i]
    ie = hdr.index('Instructions Executed')
    si = next(i for i, h in enumerate(hdr) if h.startswith('# Samples') or h == 'Warp Stall Sampling (All Samples)')
    lines = []
    fname = ''
    for r in rows:
        if r and r[0] == 'File Name':
            fname = r[1].split('/')[-1]
            continue
        if len(r) < len(hdr) or r[0] in ('', 'Line No'):
            continue
        lines.append((fname, r[0], r[1].strip(), to_int(r[ie]), to_int(r[si])))
    ti, ts = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
    print(f'# {ti} warp instructions, {ts} stall samples over {len(lines)} source lines')
    print('# by instructions executed')
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f'{l[0]}:{l[1]:>5}  inst {100 * l[3] / ti:5.1f}%  samples {100 * l[4] / ts:5.1f}%  {l[2][:140]}')
    print('# by stall samples')
    for l in sorted(lines, key=lambda l: -l[4])[:top]:
        print(f'{l[0]}:{l[1]:>5}  inst {100 * l[3] / ti:5.1f}%  samples {100 * l[4] / ts:5.1f}%  {l[2][:140]}')


if __name__ == '__main__':
    main()
