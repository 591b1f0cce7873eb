"""Join the SASS page of an ncu report (csv from `ncu -i rep --page source --csv --print-source sass`) with the line table
of the object the kernel was built from: executed warp instructions and stall samples per source line.
usage: python scripts/ncu_lines.py <sass.csv> <object.o> <kernel-regex> [top N]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

csvp, obj, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
line_of = {}
for cubin in glob.glob(d + "/*.cubin"):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur, line, stack = None, None, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if cur is None or not re.search(pat, cur):
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and line:
            line_of[int(m.group(1), 16)] = line
rows = list(csv.reader(open(csvp)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {c: i for i, c in enumerate(hdr)}
ex, smp, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
stall = collections.defaultdict(collections.Counter)
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0
base = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", 0))
    e = num(r[ci["Instructions Executed"]])
    ex[key] += e
    smp[key] += num(r[ci["Warp Stall Sampling (All Samples)"]])
    op = re.sub(r"^@!?U?P\d+\s+", "", r[ci["Source"]].strip()).split()[0].split(".")[0]
    ops[key][op] += e
    for c in stall_cols:
        stall[key][c[6:]] += num(r[ci[c]])
te, ts = sum(ex.values()), sum(smp.values())
print("total executed {:.4g}  samples {:.4g}  mapped lines {}".format(te, ts, len(line_of)))
for key, e in sorted(ex.items(), key=lambda x: -smp[x[0]])[:top]:
    print("{:5.1f}% ex {:5.1f}% smp  {}:{:<5} {} | {}".format(100 * e / te, 100 * smp[key] / max(ts, 1), key[0], key[1],
          " ".join("{}:{:.0f}%".format(k, 100 * v / max(e, 1)) for k, v in ops[key].most_common(4)),
          " ".join("{}:{:.0f}%".format(k, 100 * v / max(smp[key], 1)) for k, v in stall[key].most_common(3))))
