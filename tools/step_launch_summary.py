#!/usr/bin/env python
""" One training step out of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none [--cache-control none] --csv --log-file X.csv
python bench.py --workload W --steps 2 --warmup 3 ...`): the kernel launches between the last two AdamW launches (= one replay of the captured step),
summed per kernel. Per-launch times under ncu are serialised (and cold-cache without --cache-control none): compare SHARES, not absolutes.
Usage: step_launch_summary.py X.csv "title" > profiles/rNN_launch_summary_W.txt """
import collections
import csv
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols, data = rows[hdr], rows[hdr + 1:]
ki, vi = cols.index('Kernel Name'), cols.index('Metric Value')
marks = [i for i, r in enumerate(data) if 'adamw' in r[ki]]
assert len(marks) >= 2, 'fewer than two optimizer launches captured'
seg = data[marks[-2] + 1:marks[-1] + 1]
per = collections.OrderedDict()
for r in seg:
    name = r[ki].split('(')[0][:90]
    t = float(r[vi].replace(',', ''))
    per.setdefault(name, [0, 0.0])
    per[name][0] += 1
    per[name][1] += t
tot = sum(v[1] for v in per.values())
print(f'# {title}')
print(f'# one replay of the captured step: the {len(seg)} kernel launches between the last two AdamW launches of {len(data)} captured (ncu gpu__time_duration.sum)')
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f'{v[1] / 1e3:9.1f} us {v[0]:4d}x {v[1] / v[0] / 1e3:8.1f} us/launch {100 * v[1] / tot:5.1f}%  {k}')
print(f'total {tot / 1e3:.1f} us')
