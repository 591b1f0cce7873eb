"""Kernel-level timing probe of the fused forward (solve + read-out + masked SSE) against the two-launch path.
Usage: python scripts/kbench_sse.py [--patients N] [--D 8] [--obs 40] [--reps 5]"""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from hybrid_ode_neurips_2021_b200 import _lib as L, ops, solver  # noqa: E402
import hybrid_ode_neurips_2021_b200 as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--patients", type=int, default=1 << 20)
ap.add_argument("--D", type=int, default=8)
ap.add_argument("--obs", type=int, default=40)
ap.add_argument("--h", type=float, default=0.0625)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--lib", default=None, help="alternative build of libhode_b200.so (A/B measurements)")
args = ap.parse_args()
dev = "cuda:0"
lib = L.HodeLib(args.lib) if args.lib else L.get_lib()
B, D, obs = args.patients, args.D, args.obs
torch.manual_seed(0)
m = H.RocheODE(D, 1, 14, 1, device=dev)
y0 = torch.empty(B, D, device=dev).exponential_(100.0)
a = torch.zeros(15, B, 1, device=dev)
a[torch.randint(0, 14, (B,), device=dev), torch.arange(B, device=dev), 0] = torch.rand(B, device=dev) * 10 + 1e-3
m.set_action(a)
x = torch.randn(15, B, obs, device=dev)
mask = (torch.rand(15, B, obs, device=dev) < 0.5).float()
lin = torch.nn.Linear(D, obs).to(dev)
tt = torch.arange(0, 15.0, device=dev)
grid = solver.fixed_grid_points(tt.cpu(), args.h).to(dev)
cfg = ops.make_cfg(L.FIELD_ROCHE, D, L.RK4_38, n_dose=1, expert_grads=False, hill2=True)
pb = ops.Problem(cfg, 1, B, m.dosage, m._dose_t_f32, solver.pack_params(m, L.FIELD_ROCHE).detach()[None].contiguous(), None)


def ev(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), r


W, b = lin.weight.detach(), lin.bias.detach()
t_fused, _ = ev(lambda: ops.fixed_fwd_sse(lib, pb, y0, grid, tt, W, b, x, mask, B, want_tape=True))
t_fused_nt, _ = ev(lambda: ops.fixed_fwd_sse(lib, pb, y0, grid, tt, W, b, x, mask, B, want_tape=False, want_param_grads=False))
t_fwd, (h, tape) = ev(lambda: ops.fixed_fwd(lib, pb, y0, grid, tt, True))
t_dec, _ = ev(lambda: ops.decode_sse(lib, h, W, b, x, mask, B))
print(json.dumps({"patients": B, "D": D, "obs": obs, "fused_ms": t_fused, "fused_fwd_only_ms": t_fused_nt, "fwd_ms": t_fwd,
                  "decode_ms": t_dec, "two_launch_ms": t_fwd + t_dec}))
