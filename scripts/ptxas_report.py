"""Summarise `nvcc -Xptxas -v` output (registers, spills, stack) per kernel.  Usage: ... 2>&1 | python scripts/ptxas_report.py"""
import re
import subprocess
import sys

txt = sys.stdin.read()
cur = None
rows = {}
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
        rows[cur] = {}
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        rows[cur].update(stack=int(m.group(1)), sst=int(m.group(2)), sld=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m:
        rows[cur]["regs"] = int(m.group(1))
names = list(rows)
try:
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
except Exception:
    dem = names
for n, d in zip(names, dem):
    r = rows[n]
    d = re.sub(r"\(hode::SolveArgs, int\)|void hode::", "", d)
    print("{:>4} regs  stack {:>5}  spill st/ld {:>5}/{:<5}  {}".format(r.get("regs", -1), r.get("stack", 0), r.get("sst", 0), r.get("sld", 0), d[:110]))
