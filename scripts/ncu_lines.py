#!/usr/bin/env python
"""Per CUDA source line: executed warp instructions and stall samples of one kernel launch in an ncu report
(--import-source on, -lineinfo).  usage: ncu_lines.py REPORT KERNEL_REGEX [LAUNCH_SKIP] [MIN_PCT]"""
import csv, io, subprocess, sys
rep, regex = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name',
                      'regex:' + regex, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
fname, hdr, lines = None, None, []
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == 'File Name':
        fname = r[1].split('/')[-1]; hdr = None
    elif 'Instructions Executed' in r:
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == '-':
        g = lambda n: int(r[hdr.index(n)]) if r[hdr.index(n)].isdigit() else 0
        lines.append((fname, int(r[0]), r[1].strip(), g('# Samples'), g('Instructions Executed'), g('stall_barrier'),
                      g('stall_long_sb'), g('stall_short_sb'), g('Thread Instructions Executed')))
ts, te = sum(l[3] for l in lines) or 1, sum(l[4] for l in lines) or 1
print(f"total warp inst {te}, samples {ts}")
print(" inst%  samp%  thr/inst (barrier long_sb short_sb % of samples)  file:line  source")
for l in sorted(lines, key=lambda l: (l[0], l[1])):
    if 100 * l[4] / te >= thr or 100 * l[3] / ts >= thr:
        print(f"{100*l[4]/te:6.2f} {100*l[3]/ts:6.2f}  {l[8]/max(l[4],1):5.1f}  ({100*l[5]/ts:5.2f} {100*l[6]/ts:5.2f} {100*l[7]/ts:5.2f})  {l[0]}:{l[1]}  {l[2][:100]}")
