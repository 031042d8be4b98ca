#!/usr/bin/env python
"""Per CUDA source line: share of executed warp instructions and of stall samples for one kernel of an ncu
report captured with --import-source on (compile with -lineinfo).
usage: ncu_hot_lines.py REPORT KERNEL_REGEX [LAUNCH_SKIP] [MIN_PCT]"""
import csv, io, subprocess, sys
rep, regex = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name',
                      'regex:' + regex, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, fname, hdr = [], None, None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif 'Instructions Executed' in r and 'Line No' in r:
        hdr = r
    elif hdr and len(r) > 8 and r[0].isdigit():
        isamp, iex = hdr.index('# Samples'), hdr.index('Instructions Executed')
        ibar, ilsb, issb = hdr.index('stall_barrier'), hdr.index('stall_long_sb'), hdr.index('stall_short_sb')
        f = lambda x: int(x) if x.isdigit() else 0
        lines.append((fname, int(r[0]), r[1].strip(), f(r[isamp]), f(r[iex]), f(r[ibar]), f(r[ilsb]), f(r[issb])))
ts, te = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
print(f"total warp inst {te}, samples {ts}")
print(" inst%  samp%  (barrier long_sb short_sb %% of all samples)  file:line  source")
for l in sorted(lines, key=lambda l: (l[0], l[1])):
    if 100 * l[4] / te >= thr or 100 * l[3] / ts >= thr:
        print(f"{100*l[4]/te:6.2f} {100*l[3]/ts:6.2f}  ({100*l[5]/ts:5.2f} {100*l[6]/ts:5.2f} {100*l[7]/ts:5.2f})  {l[0]}:{l[1]}  {l[2][:110]}")
