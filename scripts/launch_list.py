#!/usr/bin/env python
"""Per-kernel times of the LAST product in an ncu `--metrics gpu__time_duration.sum --csv` log (one line per launch)."""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', '')), x['Metric Unit'], x.get('Grid Size'))
        for x in csv.DictReader(lines) if x['Metric Name'] == 'gpu__time_duration.sum']
marker = sys.argv[2] if len(sys.argv) > 2 else 'k_flop'
idx = [i for i, x in enumerate(rows) if marker in x[0]]
last = rows[idx[-1]:] if idx else rows
tot = 0
for k, v, u, g in last:
    v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    tot += v
    name = re.sub(r'\(.*', '', k).replace('<unnamed>::', '').replace('void ', '')
    print(f"{v:10.1f} us  grid={g:>14} {name}")
print(f"{tot:10.1f} us  total")
