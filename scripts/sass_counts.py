"""SASS evidence for profiles/: opcode counts of the hot kernels from the objects the shipped library was linked from
(hybrid_ode_neurips_2021_b200/csrc/build/*.o, cuobjdump -sass).  Usage: python scripts/sass_counts.py > profiles/rNN_sass_counts.txt
Shows per kernel: instructions, registers, stack (spill) bytes, and the counts that prove the instruction selection the
design relies on -- FFMA2 / FADD2 / FMUL2 (packed FP32), UBLKCP + SYNCS (TMA bulk copies + mbarrier), LDL / STL (none = no spills)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "hybrid_ode_neurips_2021_b200", "csrc", "build")
WANT = [
    ("inst_roche_d8_h1.o", r"fixed_fwd_sse_kernelINS_5RocheILi8ELb1ELb0EEELi2ELi40", "fixed_fwd_sse_kernel<Roche<8,hill2>, rk4, obs 40>  (bench step, forward)"),
    ("inst_roche_d8_h1.o", r"fixed_bwd_kernelINS_5RocheILi8ELb1ELb0EEELi2ELb0ELi1ELb1", "fixed_bwd_kernel<Roche<8,hill2>, rk4>  (bench step, reverse sweep)"),
    ("inst_roche_d8_h1.o", r"fixed_fwd_kernelINS_5RocheILi8ELb1ELb0EEELi2ELi1ELb1", "fixed_fwd_kernel<Roche<8,hill2>, rk4>"),
    ("inst_roche_d8_h1.o", r"fixed_adj_kernelINS_5RocheILi8ELb1ELb0EEELi2ELb0ELi1ELb1", "fixed_adj_kernel<Roche<8,hill2>, rk4>"),
    ("inst_roche_d12_h1.o", r"dopri5_fwd_seg_kernelINS_5RocheILi12ELb1ELb0EEELi1ELb1", "dopri5_fwd_seg_kernel<Roche<12,hill2>>  (C3 forward)"),
    ("inst_roche_d12_h1.o", r"dopri5_bwd_kernelINS_5RocheILi12ELb1ELb0EEELb0ELi1ELb1", "dopri5_bwd_kernel<Roche<12,hill2>>  (C3 reverse sweep)"),
    ("inst_roche_d12_h1.o", r"dopri5_adj_kernelINS_5RocheILi12ELb1ELb0EEELb0ELb0ELi1ELi128ELb1", "dopri5_adj_kernel<Roche<12,hill2>, batch-coupled>"),
    ("inst_roche_d6_h1.o", r"dopri5_fwd_kernelINS_5RocheILi6ELb1ELb0EEELb0ELi1ELi128ELb1", "dopri5_fwd_kernel<Roche<6,hill2>, batch-coupled>  (C1 forward)"),
    ("inst_roche_d6_h1.o", r"dopri5_bwd_kernelINS_5RocheILi6ELb1ELb0EEELb0ELi1ELb1", "dopri5_bwd_kernel<Roche<6,hill2>>  (C1 reverse sweep)"),
    ("hode_aux.o", r"decode_sse_fast_kernelILi8ELi64ELi1", "decode_sse_fast_kernel<8, 64, 1>"),
]
KEYS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "LDS", "STS", "LDG", "STG", "LDC", "LDCU", "UBLKCP", "SYNCS", "LDL", "STL",
        "BRX", "ATOMS", "RED"]
for obj, pat, title in WANT:
    path = os.path.join(OBJ, obj)
    if not os.path.exists(path):
        continue
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", path], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.search(r"Function (\S+):", ln)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", ln)
        if m and cur:
            usage[cur] = m.groups()
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        if not re.search(pat, name):
            continue
        c = collections.Counter()
        n = 0
        for ln in f.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
            if m:
                body = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
                c[body.split()[0].split(".")[0]] += 1
                n += 1
        reg, stack, _ = usage.get(name, ("?", "?", "?"))
        print("{}\n  {}\n  {} instructions ({:.1f} KB), {} registers, {} B stack\n  {}".format(
            title, name[:110], n, n * 16 / 1024, reg, stack, "  ".join("{} {}".format(k, c[k]) for k in KEYS if c[k])))
        break
