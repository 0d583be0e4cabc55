#!/usr/bin/env python
""" `ncu -i X.ncu-rep --page raw --csv` -> {kernel name: {dram_bytes (read + written, per launch, median over the captured launches), us}}: the `roofline.traffic`
figure bench.py quotes. Usage: ncu_traffic.py raw.csv [raw2.csv ...] > profiles/rNN_traffic.json """
import csv
import json
import statistics
import sys

UNIT = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out = {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi, ti = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
    per = {}
    for r in rows[2:]:
        name = r[ki].split('(')[0].replace('void ', '').replace('dcv::', '').replace('sc::', '').replace('tc::', '').strip()
        per.setdefault(name, []).append((float(r[ri]) * UNIT[units[ri]] + float(r[wi]) * UNIT[units[wi]], float(r[ti])))
    for name, v in per.items():
        out[name] = dict(dram_bytes=statistics.median(b for b, _ in v), us=statistics.median(t for _, t in v), launches=len(v), source=path.split('/')[-1])
print(json.dumps(out, indent=1, sort_keys=True))
