#!/usr/bin/env python
""" Source-level digest of an `ncu --set full --import-source on` capture: for the first kernel of `ncu -i X.ncu-rep --page source --csv`, the SASS lines that
hold most warp-stall samples, the lines that mark the warp roles (TMA / tcgen05 / mbarrier / barrier instructions, with how often they executed — e.g.
how many times each try_wait was retried) and the sample / instruction totals per 100-line bucket. This is how the round-1 findings in DESIGN.md section
3.1 were read (producer spinning on `empty`, per-unit integer divisions, MEMBAR behind fence.proxy.async, ...).

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_hotspots.py src.csv [min_share_percent] """
import csv
import sys

MARKERS = ('TRYWAIT', 'UTCHMMA', 'UTCBAR', 'UTMALDG', 'UBLKCP', 'LDGSTS', 'BAR.SYNC', 'FENCE', 'MEMBAR', 'LDTM', 'NANOSLEEP', 'ARRIVE', 'STG.E.256', 'ELECT', 'BRA.U.ANY')


def to_int(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    path = sys.argv[1]
    share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = []
            blocks.append((r[1], cur))
        elif cur is not None:
            cur.append(r)
    name, blk = blocks[0]
    hdr, data = blk[0], blk[1:]
    si, src, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    total = sum(to_int(r[si]) for r in data) or 1
    print(f'# {name[:100]}\n# {total} samples over {len(data)} SASS lines, {sum(to_int(r[ie]) for r in data)} warp instructions')
    print('# line  samples  share%  executed  instruction')
    for i, r in enumerate(data):
        s = to_int(r[si])
        if s >= total * share / 100 or any(m in r[src] for m in MARKERS):
            print(f'{i:6d} {s:8d} {100 * s / total:6.1f} {to_int(r[ie]):10d}  {r[src].strip()[:110]}')
    print('# bucket(100 lines)  samples  executed')
    for b in range(0, len(data), 100):
        seg = data[b:b + 100]
        s, e = sum(to_int(r[si]) for r in seg), sum(to_int(r[ie]) for r in seg)
        if s >= total * 0.005:
            print(f'{b:6d} {s:8d} {e:12d}')


if __name__ == '__main__':
    main()
