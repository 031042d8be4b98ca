"""Compile with -Xptxas -v and print one line per kernel: registers, spills, static smem."""
import re, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse_matrix_b200 import build as B
flt = sys.argv[1] if len(sys.argv) > 1 else ""
cmd = [B.nvcc_path(), "-Xptxas=-v", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
       "-Xcompiler", "-fPIC", "-shared", "-o", B.SO] + [os.path.join(B.CSRC, s) for s in B.SOURCES]
out = subprocess.run(cmd, capture_output=True, text=True)
txt = out.stderr + out.stdout
if out.returncode != 0:
    print(txt[-4000:]); sys.exit(1)
cur = None
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0].replace("void ", "")
        continue
    if "spill" in line and cur:
        sp = re.findall(r"(\d+) bytes spill", line)
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        sm = re.search(r"(\d+) bytes smem", line)
        if flt in cur:
            print(f"{cur:60s} regs={m.group(1):>3s} spill={'/'.join(sp)} smem={sm.group(1) if sm else 0}")
        cur = None
