"""Generate ``tests/golden/real_*.npz`` by running the REFERENCE's own real-data classes (``/root/reference/model.py``:
``RocheODEReal``, ``NeuralODEReal``, ``NeuralODEReal2nd``, ``DecoderReal``) on a synthetic ICU-shaped cohort
(``run_real.py:31-49``: obs 24, static 11, action 1, hidden int(36 * 1.2) = 43, t0 = 24), with ``oracle.odeint`` injected as
``torchdiffeq``.  Run in the build container only:  ``python -m oracle.make_golden_real``"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CPU = torch.device("cpu")


def icu_cohort(B, Z, T, obs, seed):
    g = np.random.RandomState(seed)
    y0 = (g.normal(size=(B, Z)) * 0.3).astype(np.float32)
    a = (g.uniform(size=(T, B, 1)) * (g.uniform(size=(T, B, 1)) < 0.25)).astype(np.float32)
    s = g.normal(size=(T, B, 11)).astype(np.float32)
    x = g.normal(size=(T, B, obs)).astype(np.float32)
    mask = (g.uniform(size=(T, B, obs)) < 0.5).astype(np.float32)
    return y0, a, s, x, mask


def sd_np(module, prefix):
    return {prefix + k.replace(".", "__"): v.detach().numpy().copy() for k, v in module.state_dict().items()}


def main():
    M = refload.load("model")
    T, obs, H, t0, B = 48, 24, 43, 24, 6
    for name, ode_type, Z, method in (("real_hybrid_z20_midpoint", "hybrid", 20, "midpoint"),
                                      ("real_expert_z4_midpoint", "expert", 4, "midpoint"),
                                      ("real_neural_z20_midpoint", "neural", 20, "midpoint"),
                                      ("real_2nd_z40_rk4", "2nd", 40, "rk4")):
        torch.manual_seed(11)
        y0, a, s, x, mask = icu_cohort(B, Z, T, obs, seed=Z)
        dec = M.DecoderReal(obs, Z, 1, 11, H, T, 1.0, t0=t0, method=method, ode_step_size=1.0, ode_type=ode_type, device=CPU)
        z = torch.from_numpy(y0).requires_grad_(True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x_hat, h = dec(z, torch.from_numpy(a), torch.from_numpy(s))
        # VariationalInferenceReal.loss likelihood (model.py:1247), weight = 1
        lik = torch.sum((torch.from_numpy(x)[t0:] - x_hat) ** 2 * torch.from_numpy(mask)[t0:]) / B
        lik.backward()
        # field values at a few (perturbed) times, straight from the reference class
        ts = np.array([23.0, 24.000002, 24.5, 30.999998, 31.0, 47.0], dtype=np.float32)
        yy = torch.from_numpy((np.random.RandomState(1).normal(size=(B, Z)) * 0.5).astype(np.float32))
        with torch.no_grad():
            f = np.stack([dec.ode(torch.tensor(t), yy).numpy() for t in ts])
        out = {"y0": y0, "action": a, "static": s, "x": x, "mask": mask, "h": h.detach().numpy(),
               "x_hat": x_hat.detach().numpy(), "loss": np.float32(lik.item()), "grad_y0": z.grad.numpy(), "ts": ts,
               "field_y": yy.numpy(), "field_f": f, "meta": np.array([T, obs, H, t0, B, Z])}
        out.update(sd_np(dec, "sd__"))
        for k, p in dec.named_parameters():
            out["grad__" + k.replace(".", "__")] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        np.savez(os.path.join(OUT, name + ".npz"), **out)
        print(name, "loss", lik.item(), "h range", float(h.min()), float(h.max()))


if __name__ == "__main__":
    main()
