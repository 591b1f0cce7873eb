"""Kernel-level timing probe for the dopri5 kernels (development aid), C3-like: G groups of `batch` patients, each group its
own controller (batch-coupled) or one controller per patient.  Usage:
    python scripts/kbench_dopri5.py [--groups G] [--batch 10] [--D 12] [--ctrl batch|trajectory] [--rtol 1e-7 --atol 1e-8]"""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from hybrid_ode_neurips_2021_b200 import _lib as L, ops, solver  # noqa: E402
import hybrid_ode_neurips_2021_b200 as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--groups", type=int, default=16384)
ap.add_argument("--batch", type=int, default=10)
ap.add_argument("--D", type=int, default=12)
ap.add_argument("--ctrl", default="batch")
ap.add_argument("--rtol", type=float, default=1e-7)
ap.add_argument("--atol", type=float, default=1e-8)
ap.add_argument("--cap", type=int, default=768)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--hill2", type=int, default=1)
ap.add_argument("--eg", type=int, default=0)
ap.add_argument("--lib", default=None)
ap.add_argument("--peak", type=float, default=72.1, help="FP32 FMA peak in TFLOP/s for the printed roofline fractions")
args = ap.parse_args()
dev = "cuda:0"
lib = L.get_lib() if args.lib is None else L.HodeLib(args.lib)
G, Bg, D = args.groups, args.batch, args.D
B = G * Bg
torch.manual_seed(0)
m = H.RocheODE(D, 1, 14, 1, device=dev)
y0 = torch.empty(B, D, device=dev).exponential_(100.0)
a = torch.zeros(15, B, 1, device=dev)
a[torch.randint(0, 14, (B,), device=dev), torch.arange(B, device=dev), 0] = torch.rand(B, device=dev) * 10 + 1e-3
m.set_action(a)
tt = torch.arange(0, 15.0, device=dev, dtype=torch.float64)
ctrl = L.CTRL_TRAJ if args.ctrl == "trajectory" else L.CTRL_BATCH
cfg = ops.make_cfg(L.FIELD_ROCHE, D, L.DOPRI5, n_dose=1, expert_grads=bool(args.eg), hill2=bool(args.hill2), controller=ctrl,
                   rtol=args.rtol, atol=args.atol)
pb = ops.Problem(cfg, G, Bg, m.dosage, m._dose_t_f32, solver.pack_params(m, L.FIELD_ROCHE).detach()[None].contiguous(), None)


def ev(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), r


t_f0, (h, stats, _) = ev(lambda: ops.dopri5_fwd(lib, pb, y0, tt, 0))
t_f, (h, stats, tape) = ev(lambda: ops.dopri5_fwd(lib, pb, y0, tt, args.cap))
st = stats.cpu()
assert int(st[:, 3].max()) == 0, "solver status {}".format(st[:, 3].unique())
gh = torch.randn_like(h)
t_b, _ = ev(lambda: ops.dopri5_bwd(lib, pb, tt, gh, tape, stats))
per = Bg if ctrl == L.CTRL_BATCH else 1
acc, rej = int(st[:, 0].sum()) * per, int(st[:, 1].sum()) * per
Ff = 29 + 2 * D * (D - 4) + (D - 4) + 3 + (D - 4)
fl_att, fl_acc = 6 * Ff + 64 * D, 44 * D
fwd_flops = (acc + rej) * fl_att + acc * fl_acc
bwd_flops = 3 * acc * (fl_att + fl_acc)
print(json.dumps({"roof_fwd": fwd_flops / t_f / 1e9 / args.peak, "roof_bwd": bwd_flops / t_b / 1e9 / args.peak, "groups": G, "batch": Bg, "D": D, "ctrl": args.ctrl, "rtol": args.rtol, "fwd_notape_ms": t_f0, "fwd_ms": t_f, "bwd_ms": t_b,
                  "acc_per_ctrl": float(st[:, 0].float().mean()), "rej_per_ctrl": float(st[:, 1].float().mean()),
                  "max_acc": int(st[:, 0].max()), "traj_attempts": acc + rej,
                  "fwd_Gattempts_s": (acc + rej) / t_f / 1e6, "fwdbwd_Gattempts_s": (acc + rej) / (t_f + t_b) / 1e6,
                  "bwd_Gacc_s": acc / t_b / 1e6}))
