"""In-tree build of ``libhode_b200.so`` for sm_100a with nvcc (cross-compiles without a GPU).

``python -m hybrid_ode_neurips_2021_b200.build [--force]``.  Objects go to ``csrc/build/`` (git-ignored), the shared
library next to this file (git-ignored, but it travels to the GPU box with the repo snapshot).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
OUT = os.path.join(HERE, "libhode_b200.so")
ROOT = os.path.dirname(HERE)

ROCHE_DIMS = (4, 6, 8, 12)
NEURAL_DIMS = (4, 6, 8, 12)
# real-data fields: (hode_field, latent width) -- RocheODEReal 4 / 20, NeuralODEReal 4 / 20, NeuralODEReal2nd 8 / 40
REAL_UNITS = ((2, 4), (2, 20), (3, 4), (3, 20), (4, 8), (4, 40))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libhode_b200.so cannot be built")
    return exe


def _units():
    """(object name, source, extra defines)"""
    units = [("hode_api.o", "hode_api.cu", []), ("hode_aux.o", "hode_aux.cu", []), ("hode_eval.o", "hode_eval.cu", []), ("hode_real.o", "hode_real.cu", [])]
    for d in ROCHE_DIMS:
        for hill2 in (0, 1, 2):  # 2 = the ablation field
            units.append(("inst_roche_d{}_h{}.o".format(d, hill2), "inst_roche.cu",
                          ["-DHODE_INST_D={}".format(d), "-DHODE_INST_HILL2={}".format(hill2)]))
    for field, z in REAL_UNITS:
        units.append(("inst_real_f{}_z{}.o".format(field, z), "inst_real.cu",
                      ["-DHODE_REAL_FIELD={}".format(field), "-DHODE_REAL_Z={}".format(z)]))
    if os.path.exists(os.path.join(CSRC, "inst_neural.cu")):
        for d in NEURAL_DIMS:
            units.append(("inst_neural_d{}.o".format(d), "inst_neural.cu", ["-DHODE_INST_D={}".format(d)]))
    return units


def _source_digest(src, extra) -> str:
    """Digest of one unit's inputs: its own .cu, every header of csrc/ (conservative) and include/hode.h."""
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name == src or name.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    with open(os.path.join(ROOT, "include", "hode.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    return h.hexdigest()


def _compile(obj, src, defs, extra):
    stamp = os.path.join(OBJ, obj + ".stamp")
    dig = _source_digest(src, list(defs) + list(extra))
    out = os.path.join(OBJ, obj)
    if os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, "cached", ""
    cmd = [nvcc()] + NVCC_FLAGS + list(defs) + list(extra) + ["-c", os.path.join(CSRC, src), "-o", out]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for {}:\n{}\n{}".format(obj, " ".join(cmd), p.stderr[-4000:]))
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, "built", p.stderr


def build(force: bool = False, extra_flags=(), verbose: bool = True, jobs: int | None = None) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for n in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, n))
    units = _units()
    jobs = jobs or min(len(units), os.cpu_count() or 4)
    rebuilt = False
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        futs = [ex.submit(_compile, o, s, d, extra_flags) for o, s, d in units]
        for f in futs:
            obj, state, log = f.result()
            rebuilt |= state == "built"
            if verbose:
                print("[hode build] {:<22} {}".format(obj, state), flush=True)
    if rebuilt or not os.path.exists(OUT):
        cmd = [nvcc(), "-shared", "-o", OUT] + [os.path.join(OBJ, o) for o, _, _ in units] + ["-lcudart"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stderr[-4000:])
        if verbose:
            print("[hode build] linked", OUT, flush=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
