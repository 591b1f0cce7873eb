"""Decode + masked-SSE kernel timing (development aid).  Usage: python scripts/kbench_decode.py --D 8 --obs 40 --patients N"""
import argparse, json, sys
import torch
sys.path.insert(0, ".")
from hybrid_ode_neurips_2021_b200 import _lib as L, ops  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--patients", type=int, default=1 << 20)
ap.add_argument("--D", type=int, default=8)
ap.add_argument("--obs", type=int, default=40)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = "cuda:0"
lib = L.get_lib()
B, D, obs = a.patients, a.D, a.obs
torch.manual_seed(0)
h = torch.randn(15, B, D, device=dev)
x = torch.randn(15, B, obs, device=dev)
mask = (torch.rand(15, B, obs, device=dev) < 0.5).float()
lin = torch.nn.Linear(D, obs).to(dev)
fn = lambda: ops.decode_sse(lib, h, lin.weight.detach(), lin.bias.detach(), x, mask, B)
fn(); torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
by = 15 * B * (2 * obs + 2 * D) * 4
import os
print(json.dumps({"variant": os.environ.get("HODE_DECODE_VARIANT", "default"), "D": D, "obs": obs, "B": B, "ms": min(ts), "GBs": by / min(ts) / 1e6}))
