import csv, collections, sys
path = sys.argv[1]; tail = int(sys.argv[2]) if len(sys.argv) > 2 else 150
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]; data = rows[hdr+1:]
ki, vi = cols.index('Kernel Name'), cols.index('Metric Value')
per = collections.OrderedDict()
for r in data[-tail:]:
    name = r[ki].split('(')[0][:80]
    t = float(r[vi].replace(',', ''))
    per.setdefault(name, [0, 0.0]); per[name][0] += 1; per[name][1] += t
tot = sum(v[1] for v in per.values())
print(f'# last {tail} launches of {len(data)} captured (ncu gpu__time_duration.sum; cold-cache, serialised: compare shares)')
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f'{v[1]/1e3:9.1f} us {v[0]:4d}x {v[1]/v[0]/1e3:8.1f} us/launch {100*v[1]/tot:5.1f}%  {k}')
print(f'total {tot/1e3:.1f} us')
