"""Aggregate the source page of an .ncu-rep per CUDA source line (run where ncu is installed; needs -lineinfo and
--import-source on): python scripts/ncu_source.py <file.ncu-rep> [top N]
Prints, per source line, executed warp-level instructions and the sampled stall counts, largest first."""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
recs = []
for r in rows:
    if hdr is None:
        if "Source" in r and any("Instructions Executed" in c for c in r):
            hdr = r
        continue
    if len(r) == len(hdr):
        recs.append(r)
if hdr is None:
    print(out[:2000])
    sys.exit(1)
ci = {c: i for i, c in enumerate(hdr)}
def col(name):
    for c, i in ci.items():
        if c.strip() == name:
            return i
    return None
i_src, i_ex, i_smp = col("Source"), col("Instructions Executed"), col("Warp Stall Sampling (All Samples)") or col("Warp Stall Sampling (All Cycles)")
i_file = col("File") if col("File") is not None else None
def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0
tot_ex = sum(num(r[i_ex]) for r in recs)
tot_smp = sum(num(r[i_smp]) for r in recs) if i_smp is not None else 0
print("columns:", [c for c in hdr][:12])
print("total executed {:.4g} sampled {:.4g}".format(tot_ex, tot_smp))
recs.sort(key=lambda r: -num(r[i_ex]))
for n, r in enumerate(recs[:top]):
    print("{:6.2f}% ex {:6.2f}% smp | {}".format(100 * num(r[i_ex]) / max(tot_ex, 1), 100 * num(r[i_smp]) / max(tot_smp, 1) if i_smp is not None else 0,
                                              (r[i_src] or "").strip()[:150]))
