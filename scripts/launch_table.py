"""Print one line per kernel launch from an ncu --csv log with time / grid / instructions."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; ki = h.index('Kernel Name'); mi = h.index('Metric Name'); vi = h.index('Metric Value'); ii = h.index('ID')
d = collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r) > vi:
        k = (r[ii], r[ki].split('(')[0].replace('void <unnamed>::','').replace('<unnamed>::',''))
        d.setdefault(k, {})[r[mi]] = float(r[vi].replace(',',''))
names = []
for (i, name), m in d.items():
    if name in names and len(sys.argv) < 3: break
    names.append(name)
    print(f"{i:>4s} {name[:48]:48s} t_us={m.get('gpu__time_duration.sum',0)/1e3:10.1f} grid={int(m.get('launch__grid_size',0)):8d} inst={m.get('smsp__inst_executed.sum',0)/1e6:9.1f}M")
