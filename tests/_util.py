"""Shared helpers for the parity tests: synthetic cohorts with the reference generator's distributions
(dataloader.py:200-266), weight exchange between the oracle modules and the drop-ins, norm-wise error metrics."""
import numpy as np
import torch

from oracle import fields as OF

EXPERT_NAMES = OF.EXPERT_NAMES


def make_cohort(B, D, T=15, obs=20, seed=0, dose_max=10.0, n_dose=1):
    g = np.random.RandomState(seed)
    y0 = g.exponential(scale=0.01, size=(B, D)).astype(np.float32)
    a = np.zeros((T, B, 1), dtype=np.float32)
    for b in range(B):
        days = g.choice(T - 1, size=n_dose, replace=False)
        a[days, b, 0] = g.uniform(0.05, dose_max)  # one amount per patient (model.py:497 takes the max)
    x = g.normal(size=(T, B, obs)).astype(np.float32)
    mask = (g.uniform(size=(T, B, obs)) < 0.5).astype(np.float32)
    return torch.from_numpy(y0), torch.from_numpy(a), torch.from_numpy(x), torch.from_numpy(mask)


def oracle_roche(D, seed=0, perturb_scalars=False):
    torch.manual_seed(seed)
    ode = OF.OracleRocheODE(D)
    if perturb_scalars:
        g = np.random.RandomState(seed + 1)
        with torch.no_grad():
            for n in EXPERT_NAMES:
                if n in ("HillCure", "HillPatho"):
                    continue
                getattr(ode, n).mul_(float(g.uniform(0.8, 1.2)))
    return ode


def relerr(a, b):
    """norm-wise ||a-b||_inf / ||b||_inf"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.abs().max().item()
    return ((a - b).abs().max().item() / den) if den > 0 else (a - b).abs().max().item()


def nan_pattern_equal(a, b):
    return bool((torch.isnan(a.cpu()) == torch.isnan(b.cpu())).all())


def grads_of(module):
    return {n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in module.named_parameters()}


def check(name, value, gate):
    """Assert ``value <= gate`` and log the measured error next to its gate (``gpurun_out/parity_report.jsonl`` when that
    directory exists): the report is how the gates are kept close to what the hardware actually delivers."""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    rec = {"check": name, "value": float(value), "gate": float(gate), "ok": bool(value <= gate)}
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    print("[parity] {:<70} {:.3e}  (gate {:.1e})".format(name, float(value), float(gate)))
    assert value <= gate, rec
