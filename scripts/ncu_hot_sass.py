"""Hot SASS of one kernel from an ncu report's source page: executions per row and stall samples."""
import csv, subprocess, sys, io
rep, regex, rows_n = sys.argv[1], sys.argv[2], float(sys.argv[3])
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 20
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + regex, '--launch-skip', skip,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Address' in r and 'Source' in r)
h = rows[hi]
isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
data = []
for r in rows[hi + 1:]:
    if len(r) > iex and r[iex].isdigit():
        data.append((r[isrc].strip(), int(r[isamp] or 0), int(r[iex])))
tot_s = sum(d[1] for d in data); tot_e = sum(d[2] for d in data)
print(rows[0][1][:100] if rows[0] else '', "| total inst/row", round(tot_e / rows_n, 1), "sass", len(data))
for i, (s, sa, ex) in enumerate(data):
    if ex > rows_n * thr or sa > tot_s * 0.015:
        print(f"{i:4d} {ex/rows_n:8.1f}x  samp={100*sa/max(1,tot_s):5.2f}%  {s[:100]}")
