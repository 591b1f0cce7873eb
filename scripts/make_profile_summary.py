"""Turn one gpurun_out/prof_<tag>/ directory (scripts/profile.sh) into the committed text evidence under profiles/:
    python scripts/make_profile_summary.py <tag> [<round label>]
writes profiles/<label>_launches.txt (per-kernel share of the step from the ncu launch list), profiles/<label>_<kernel>.txt
(key `ncu --set full` counters per hot kernel) and copies the plain bench line."""
import collections
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else tag
src = os.path.join(ROOT, "gpurun_out", "prof_" + tag)
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

# ---- launch list -----------------------------------------------------------------------------------------------------
rows = []
with open(os.path.join(src, "launches.csv")) as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)  # -> us
    rows.append((re.sub(r"\(.*", "", r[ki]), v))
agg = collections.OrderedDict()
for k, v in rows:
    c, t = agg.get(k, (0, 0.0))
    agg[k] = (c + 1, t + v)
tot = sum(t for _, t in agg.values())
with open(os.path.join(dst, label + "_launches.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none   python bench.py --steps 2 --warmup 3 --cpu-patients 512\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes. {} launches, {:.1f} ms total\n".format(len(rows), tot / 1e3))
    f.write("{:>7} {:>12} {:>12} {:>7}  kernel\n".format("count", "total_us", "mean_us", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("{:>7} {:>12.1f} {:>12.1f} {:>6.1f}%  {}\n".format(c, t, t / c, 100 * t / tot, k[:140]))
    # The list mixes full-cohort launches (the device-resident step) with the 1/8-size launches of the end-to-end phase and
    # the extra forward launches of the forward-only timing.  Shares of ONE device-resident step: mean duration of the
    # full-size launches of each hode kernel (a launch counts as full-size above half of that kernel's longest launch).
    per = collections.OrderedDict()
    for k, v in rows:
        if "hode::" in k:
            per.setdefault(k, []).append(v)
    full = {k: [v for v in vs if v > 0.5 * max(vs)] for k, vs in per.items()}
    step = {k: sum(v) / len(v) for k, v in full.items() if v and "prep_params" not in k and "ffma_probe" not in k
            and "adj_kernel" not in k and "dopri5" not in k and "crps" not in k}
    stot = sum(step.values())
    f.write("\n# one device-resident step = one full-size launch of each kernel below ({:.2f} ms):\n".format(stot / 1e3))
    for k, v in sorted(step.items(), key=lambda kv: -kv[1]):
        f.write("{:>7} {:>12.1f} {:>12} {:>6.1f}%  {}\n".format(len(full[k]), v, "(mean us)", 100 * v / stot, k[:140]))

# ---- per-kernel full captures ----------------------------------------------------------------------------------------
names = sorted(os.listdir(src))
for name in names:
    if name.endswith(".summary.txt"):  # summarised on the GPU box (scripts/profile.sh)
        out = open(os.path.join(src, name)).read()
        name = name[:-12] + ".ncu-rep"
    elif name.endswith(".ncu-rep") and name[:-8] + ".summary.txt" not in names:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), os.path.join(src, name)],
                             capture_output=True, text=True).stdout
    else:
        continue
    keep = [ln for ln in out.splitlines() if not re.search(r"occupancy_per_|fp16|_adu|_cbu|membar|sleeping|tex_throttle|misc_per|drain|syslts", ln)]
    with open(os.path.join(dst, "{}_{}.txt".format(label, name[:-8])), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:{} (bench.py --patients 131072; one launch)\n".format(name[:-8]))
        f.write("\n".join(keep) + "\n")
for name in ("plain.json", "plain_small.json"):
    if os.path.exists(os.path.join(src, name)):
        shutil.copy(os.path.join(src, name), os.path.join(dst, "{}_bench_{}".format(label, name)))
print(sorted(os.listdir(dst)))
