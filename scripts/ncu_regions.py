import csv, io, subprocess, sys
rep, regex, skip = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name',
                      'regex:' + regex, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, fname, hdr = [], None, None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]
    elif 'Instructions Executed' in r and 'Line No' in r: hdr = r
    elif hdr and len(r) > 8 and r[0].isdigit():
        isamp, iex = hdr.index('# Samples'), hdr.index('Instructions Executed')
        f = lambda x: int(x) if x.isdigit() else 0
        lines.append((fname, int(r[0]), f(r[isamp]), f(r[iex])))
ts, te = sum(l[2] for l in lines), sum(l[3] for l in lines)
regions = {
 'probe_insert (rowhash 70-95)': lambda f,l: f=='rowhash.cuh' and 70<=l<=95,
 'load_chunk/locate (rowhash 96-140)': lambda f,l: f=='rowhash.cuh' and 96<=l<=140,
 'bucket sorts helpers (rowhash 314-407)': lambda f,l: f=='rowhash.cuh' and 314<=l<=407,
 'setup+init (456-486)': lambda f,l: f=='rowhash.cuh' and 456<=l<=486,
 'accumulate loop (487-558)': lambda f,l: f=='rowhash.cuh' and 487<=l<=558,
 'drain count+scan (596-634)': lambda f,l: f=='rowhash.cuh' and 596<=l<=634,
 'drain scatter (635-645)': lambda f,l: f=='rowhash.cuh' and 635<=l<=645,
 'drain bucket loop (646-663)': lambda f,l: f=='rowhash.cuh' and 646<=l<=663,
 'fallback (664-698)': lambda f,l: f=='rowhash.cuh' and 664<=l<=698,
 'other files': lambda f,l: f!='rowhash.cuh',
}
for name, pred in regions.items():
    s = sum(l[2] for l in lines if pred(l[0], l[1])); e = sum(l[3] for l in lines if pred(l[0], l[1]))
    print(f"{100*e/te:6.2f}% inst {100*s/ts:6.2f}% samples  {name}")
