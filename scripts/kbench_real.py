"""DecoderReal timing at the ICU shape (run_real.py: T = 120 hourly steps, t0 = 24, obs 24, static 11, hidden 43, batch 100)."""
import sys, time, json
import torch
sys.path.insert(0, ".")
import hybrid_ode_neurips_2021_b200 as H
dev = "cuda:0"
T, obs, Hd, t0 = 120, 24, 43, 24
for ode_type, Z, method, B in (("hybrid", 20, "midpoint", 100), ("neural", 20, "midpoint", 100), ("2nd", 40, "rk4", 100), ("expert", 4, "midpoint", 100),
                               ("hybrid", 20, "midpoint", 16384)):
    torch.manual_seed(0)
    dec = H.DecoderReal(obs, Z, 1, 11, Hd, T, 1.0, t0=t0, method=method, ode_step_size=1.0, ode_type=ode_type, device=dev)
    y0 = (torch.randn(B, Z, device=dev) * 0.1).requires_grad_(True)
    a = torch.rand(T, B, 1, device=dev) * (torch.rand(T, B, 1, device=dev) < 0.25).float()
    s = torch.randn(T, B, 11, device=dev)
    x = torch.randn(T, B, obs, device=dev); m = (torch.rand(T, B, obs, device=dev) < 0.5).float()
    def step():
        dec.zero_grad(); y0.grad = None
        xh, h = dec(y0, a, s)
        loss = torch.sum((x[t0:] - xh) ** 2 * m[t0:]) / B
        loss.backward()
        return loss
    step(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t_ = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t_)
    print(json.dumps({"ode_type": ode_type, "Z": Z, "method": method, "B": B, "fwd_bwd_ms": min(ts) * 1e3,
                      "traj_steps_per_s": B * (T - t0) / min(ts)}))
