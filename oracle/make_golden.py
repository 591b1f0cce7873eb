"""Generate ``tests/golden/*.npz`` by running the REFERENCE's own ``model.py`` (imported unmodified from
``/root/reference`` with ``oracle.odeint`` injected as ``torchdiffeq``).  Run in the build container only:

    python -m oracle.make_golden

The fixtures pin (a) the vector fields / dose schedules / decoder read-out / masked SSE bit-for-bit to the reference's
code, and (b) whole solves + losses + gradients to "reference model.py + restated odeint" on the same inputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402
from oracle import odeint as OI  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CPU = torch.device("cpu")


def cohort(B, D, T=15, obs=20, seed=0, n_dose=1):
    g = np.random.RandomState(seed)
    y0 = g.exponential(scale=0.01, size=(B, D)).astype(np.float32)
    a = np.zeros((T, B, 1), dtype=np.float32)
    for b in range(B):
        days = g.choice(T - 1, size=n_dose, replace=False)
        a[days, b, 0] = g.uniform(0.05, 10.0)
    x = g.normal(size=(T, B, obs)).astype(np.float32)
    mask = (g.uniform(size=(T, B, obs)) < 0.5).astype(np.float32)
    return y0, a, x, mask


def sd_np(module, prefix):
    return {prefix + k.replace(".", "__"): v.detach().numpy().copy() for k, v in module.state_dict().items()}


def main():
    M = refload.load("model")
    os.makedirs(OUT, exist_ok=True)

    # ---- vector fields -------------------------------------------------------------------------------------
    for D in (4, 6, 8, 12):
        torch.manual_seed(100 + D)
        y0, a, _, _ = cohort(9, D, seed=D)
        g = np.random.RandomState(D)
        y = torch.from_numpy((g.normal(size=(9, D)) * 0.5).astype(np.float32))  # includes negative states
        ts = np.array([0.0, 2.9999998, 3.0, 4.5, 7.25, 13.999999], dtype=np.float32)
        out = {"y": y.numpy(), "action": a, "ts": ts}
        ode = M.RocheODE(D, 1, 14, 1, device=CPU)
        with torch.no_grad():
            for n in ("k_dexa", "k_disprog", "k_immune_off", "kel"):
                getattr(ode, n).mul_(float(g.uniform(0.8, 1.2)))
        ode.set_action(torch.from_numpy(a))
        out.update(sd_np(ode, "sd__"))
        out["times"] = ode.times.numpy()
        out["dosage"] = ode.dosage.numpy()
        with torch.no_grad():
            out["f"] = np.stack([ode(torch.tensor(t), y).numpy() for t in ts])
            out["dose"] = np.stack([ode.dose_at_time(torch.tensor(t)).numpy() for t in ts])
        np.savez(os.path.join(OUT, "roche_field_d{}.npz".format(D)), **out)

        node = M.NeuralODE(D, 1, 14, 1, device=CPU)
        node.set_action(torch.from_numpy(a))
        out = {"y": y.numpy(), "action": a, "ts": np.array([0.0, 2.9999998, 3.0, 5.0, 7.25], dtype=np.float32)}
        out.update(sd_np(node, "sd__"))
        with torch.no_grad():
            out["f"] = np.stack([node(torch.tensor(t), y).numpy() for t in out["ts"]])
            out["dose"] = np.stack([node.dose_at_time(torch.tensor(t)).numpy() for t in out["ts"]])
        np.savez(os.path.join(OUT, "neural_field_d{}.npz".format(D)), **out)

    # ---- whole decoder solves + loss + gradients (reference decoder + restated odeint) ----------------------
    cases = [
        ("hybrid_d6_dopri5", dict(D=6, obs=20, B=10, roche=True, method="dopri5")),
        ("hybrid_d12_dopri5", dict(D=12, obs=80, B=10, roche=True, method="dopri5")),
        ("expert_d4_dopri5", dict(D=4, obs=20, B=10, roche=True, method="dopri5")),
        ("neural_d6_dopri5", dict(D=6, obs=20, B=10, roche=False, method="dopri5")),
    ]
    for name, c in cases:
        torch.manual_seed(7)
        D, obs, B = c["D"], c["obs"], c["B"]
        y0, a, x, mask = cohort(B, D, obs=obs, seed=40 + D)
        dec = M.RocheExpertDecoder(obs, D, 1, 14, 1, roche=c["roche"], method=c["method"], device=CPU)
        z = torch.from_numpy(y0).requires_grad_(True)
        # gradients with the first step held constant (what the CUDA path implements)
        import torchdiffeq as tde
        orig = tde.odeint

        def no_first_step_grad(func, y, t, **kw):
            kw["options"] = dict(kw.get("options") or {}, differentiable_first_step=False)
            return orig(func, y, t, **kw)

        M.dto = no_first_step_grad
        try:
            x_hat, h = dec(z, torch.from_numpy(a))
        finally:
            M.dto = orig
        lik = torch.sum((torch.from_numpy(x) - x_hat) ** 2 * torch.from_numpy(mask)) / x.shape[1]
        lik.backward()
        out = {"y0": y0, "action": a, "x": x, "mask": mask, "h": h.detach().numpy(), "x_hat": x_hat.detach().numpy(),
               "loss": np.float32(lik.item()), "grad_y0": z.grad.numpy()}
        out.update(sd_np(dec, "sd__"))
        for k, p in dec.named_parameters():
            if p.grad is not None:
                out["grad__" + k.replace(".", "__")] = p.grad.numpy()
        np.savez(os.path.join(OUT, name + ".npz"), **out)
        print(name, "loss", lik.item())

    # ---- fixed-step sweep shape (config 2): dim-8 RK4 3/8 with step 1/16 ------------------------------------
    torch.manual_seed(8)
    D, B = 8, 64
    y0, a, _, _ = cohort(B, D, seed=48)
    ode = M.RocheODE(D, 1, 14, 1, device=CPU)
    ode.set_action(torch.from_numpy(a))
    with torch.no_grad():
        h = OI.odeint(ode, torch.from_numpy(y0), torch.arange(0, 15.0), method="rk4", options={"step_size": 0.0625})
    out = {"y0": y0, "action": a, "h": h.numpy()}
    out.update(sd_np(ode, "sd__"))
    np.savez(os.path.join(OUT, "hybrid_d8_rk4_h16.npz"), **out)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
