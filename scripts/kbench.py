"""Kernel-level timing probe (development aid): times the solver / decode kernels alone through the C ABI with data
generated on the GPU.  Usage: python scripts/kbench.py [--patients N] [--D 8] [--method rk4] [--what fwd,bwd,dec]"""
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
ap.add_argument("--method", default="rk4")
ap.add_argument("--eg", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--lib", default=None)
ap.add_argument("--hill2", type=int, default=1)
ap.add_argument("--field", default="roche")
args = ap.parse_args()
dev = "cuda:0"
lib = L.get_lib() if args.lib is None else L.HodeLib(args.lib)
B, D, obs = args.patients, args.D, args.obs
torch.manual_seed(0)
FIELD = L.FIELD_ROCHE if args.field == "roche" else L.FIELD_NEURAL
m = H.RocheODE(D, 1, 14, 1, device=dev) if args.field == "roche" else H.NeuralODE(D, 1, 14, 1, device=dev)
y0 = torch.empty(B, D, device=dev).exponential_(100.0)
a = torch.zeros(15, B, 1, device=dev)
a[torch.randint(0, 14, (B,), device=dev), torch.arange(B, device=dev), 0] = torch.rand(B, device=dev) * 10 + 1e-3
m.set_action(a)
x = torch.randn(15, B, obs, device=dev)
mask = (torch.rand(15, B, obs, device=dev) < 0.5).float()
lin = torch.nn.Linear(D, obs).to(dev)
tt = torch.arange(0, 15.0, device=dev)
grid = solver.fixed_grid_points(tt.cpu(), args.h).to(dev)
cfg = ops.make_cfg(FIELD, D, L.METHODS[args.method], n_dose=1, expert_grads=bool(args.eg), hill2=bool(args.hill2) and args.field == "roche")
pb = ops.Problem(cfg, 1, B, m.dosage, m._dose_t_f32, solver.pack_params(m, FIELD).detach()[None].contiguous(), None)


def ev(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), r


t_fwd0, _ = ev(lambda: ops.fixed_fwd(lib, pb, y0, grid, tt, False))
t_fwd, (h, tape) = ev(lambda: ops.fixed_fwd(lib, pb, y0, grid, tt, True))
t_dec, (loss, gh, gw, gb) = ev(lambda: ops.decode_sse(lib, h, lin.weight.detach(), lin.bias.detach(), x, mask, B))
t_bwd, _ = ev(lambda: ops.fixed_bwd(lib, pb, grid, tt, gh, tape))
n = grid.numel() - 1
adj_grid, adj_count = solver.adjoint_grid_points(tt.cpu(), args.h)
adj_grid, adj_count = adj_grid.to(dev), adj_count.to(dev)
t_adj, (gy0_a, gp_a) = ev(lambda: ops.fixed_adjoint(lib, pb, adj_grid, adj_count, h, gh))
gy0_d, gp_d = ops.fixed_bwd(lib, pb, grid, tt, gh, tape)
gap = ((gy0_a - gy0_d).abs().max() / gy0_d.abs().max()).item()
print(json.dumps({"B": B, "D": D, "method": args.method, "steps": n, "fwd_notape_ms": t_fwd0, "fwd_ms": t_fwd, "dec_ms": t_dec,
                  "bwd_ms": t_bwd, "adjoint_ms": t_adj, "adjoint_vs_discrete_dy0": gap, "fwd_Gsteps": B * n / t_fwd / 1e6,
                  "bwd_Gsteps": B * n / t_bwd / 1e6, "adjoint_Gsteps": B * n / t_adj / 1e6}))
