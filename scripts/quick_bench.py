"""Ad-hoc timing probe used during development (not the contract bench; see bench.py)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import hybrid_ode_neurips_2021_b200 as H  # noqa: E402

dev = "cuda:0"


def timeit(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def fixed_case(B, D, h, method="rk4", bwd=False):
    torch.manual_seed(0)
    m = H.RocheODE(D, 1, 14, 1, device=dev)
    y0 = torch.distributions.Exponential(100.0).sample((B, D)).to(dev)
    a = torch.zeros(15, B, 1, device=dev)
    day = torch.randint(0, 14, (B,), device=dev)
    a[day, torch.arange(B, device=dev), 0] = torch.rand(B, device=dev) * 10
    m.set_action(a)
    t = torch.arange(0, 15.0, device=dev)
    steps = int(14 / h)
    if not bwd:
        def fn():
            with torch.no_grad():
                return H.odeint(m, y0, t, method=method, options={"step_size": h})
    else:
        y0g = y0.clone().requires_grad_(True)
        def fn():
            out = H.odeint(m, y0g, t, method=method, options={"step_size": h, "expert_grads": False})
            out.backward(torch.ones_like(out))
    best, avg = timeit(fn)
    print(json.dumps({"case": "fixed", "B": B, "D": D, "h": h, "bwd": bwd, "ms_best": best, "ms_avg": avg,
                      "traj_steps_per_s": B * steps / (best * 1e-3)}), flush=True)


def dopri_case(B, G, D, rtol, atol, ctrl="batch", bwd=True):
    torch.manual_seed(0)
    m = H.RocheODE(D, 1, 14, 1, device=dev)
    N = B * G
    y0 = torch.distributions.Exponential(100.0).sample((N, D)).to(dev)
    a = torch.zeros(15, N, 1, device=dev)
    day = torch.randint(0, 14, (N,), device=dev)
    a[day, torch.arange(N, device=dev), 0] = torch.rand(N, device=dev) * 10
    m.set_action(a)
    t = torch.arange(0, 15.0, device=dev)
    y0g = y0.clone().requires_grad_(bwd)
    opts = {"n_groups": G, "controller": ctrl, "expert_grads": False, "tape_capacity": 1024}
    def fn():
        out = H.odeint(m, y0g, t, method="dopri5", rtol=rtol, atol=atol, options=opts)
        if bwd:
            out.backward(torch.ones_like(out))
    best, avg = timeit(fn)
    info = H.last_solve_info()
    att = info.attempts_total * (B if ctrl == "batch" else 1)
    print(json.dumps({"case": "dopri5", "B": B, "G": G, "D": D, "rtol": rtol, "ctrl": ctrl, "bwd": bwd, "ms_best": best,
                      "traj_attempts": att, "traj_steps_per_s": att / (best * 1e-3),
                      "acc_mean": float(info.accepted.float().mean()), "rej_mean": float(info.rejected.float().mean())}), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    fixed_case(1 << 20, 8, 0.0625)
    fixed_case(1 << 20, 8, 0.0625, bwd=True)
    fixed_case(1 << 18, 6, 0.0625)
    fixed_case(1 << 18, 12, 0.0625)
    fixed_case(1 << 18, 12, 0.0625, bwd=True)
    dopri_case(50, 1, 6, 1e-7, 1e-8)
    dopri_case(50, 1024, 6, 1e-7, 1e-8)
    dopri_case(10, 4096, 12, 1e-7, 1e-8)
    dopri_case(1, 1 << 16, 8, 1e-7, 1e-8, ctrl="trajectory")
    dopri_case(1, 1 << 16, 8, 1e-7, 1e-8, ctrl="trajectory", bwd=False)
    print("wall", time.time() - t0)
