"""Full-size checks (BASELINE configs[1]: 2^20 patients, dim-8 hybrid field, RK4 3/8-rule h = 1/16 over 14 days) through
size-independent properties, so that the cohort size the bench is quoted on is tied to the oracle:

* trajectory independence: row i of the full solve (values AND dL/dy0) is BIT-IDENTICAL to the same patient solved inside a
  random 4 096-patient subset (one thread per trajectory, no cross-trajectory arithmetic);
* the subset's first patients agree with the CPU oracle (trajectories 1e-5, dL/dy0 2e-5, norm-wise);
* gradient additivity ("checksum of checksums"): the parameter gradient of the whole cohort equals the sum over 8 contiguous
  shards (what the multi-GPU path relies on) to 5e-4 norm-wise (float32 atomics, different reduction trees).

The same body runs at a small size through the host emulation in the CPU suite (debugging aid for the test itself)."""
import os
import subprocess

import pytest
import torch

from hybrid_ode_neurips_2021_b200 import _lib as L
from hybrid_ode_neurips_2021_b200 import ops
from oracle import odeint as OI

from _util import EXPERT_NAMES, oracle_roche, relerr

HS_DIR = os.path.join(os.path.dirname(__file__), "hostsim")
D, STEP = 8, 0.0625


def run_properties(lib, dev, B, n_subset, n_shards, n_oracle):
    o = oracle_roche(D, 1, True)
    params = torch.cat([getattr(o, n).detach().reshape(1) for n in EXPERT_NAMES]
                       + [o.ml_net[0].weight.detach().reshape(-1), o.ml_net[0].bias.detach().reshape(-1)]).float()[None]
    params = params.contiguous().to(dev)
    g = torch.Generator(device=dev).manual_seed(123)
    y0 = torch.empty(B, D, device=dev).exponential_(100.0, generator=g)
    day = torch.randint(0, 14, (B,), device=dev, generator=g)
    amt = torch.rand(B, device=dev, generator=g) * 10.0 + 1e-3
    dose_t = day.to(torch.float32)[:, None].contiguous()
    t = torch.arange(0, 15.0)
    grid = OI.fixed_grid_points(t, STEP).contiguous().to(dev)
    tt = t.to(dev)
    W = torch.randn(15, B, D, device=dev, generator=g)
    cfg = ops.make_cfg(L.FIELD_ROCHE, D, L.RK4_38, n_dose=1, hill2=True, expert_grads=False)

    def solve(sel):
        pb = ops.Problem(cfg, 1, int(sel.numel()), amt[sel].contiguous(), dose_t[sel].contiguous(), params, None)
        h, tape = ops.fixed_fwd(lib, pb, y0[sel].contiguous(), grid, tt, True)
        gy0, gp = ops.fixed_bwd(lib, pb, grid, tt, W[:, sel].contiguous(), tape)
        return h, gy0, gp[0]

    everyone = torch.arange(B, device=dev)
    h, gy0, gp = solve(everyone)
    assert h.shape == (15, B, D) and bool(torch.isfinite(h).all()) and bool(torch.isfinite(gy0).all())
    assert torch.equal(h[0], y0)
    # trajectory independence, bit for bit
    sub = torch.randperm(B, device=dev, generator=g)[:n_subset]
    h_s, gy0_s, _ = solve(sub)
    assert torch.equal(h[:, sub], h_s)
    assert torch.equal(gy0[sub], gy0_s)
    # the subset against the CPU oracle
    sel = sub[:n_oracle].cpu()
    a = torch.zeros(15, n_oracle, 1)
    a[day.cpu()[sel], torch.arange(n_oracle), 0] = amt.cpu()[sel]
    o.set_action(a)
    z = y0.cpu()[sel].clone().requires_grad_(True)
    ref = OI.odeint(o, z, t, method="rk4", options={"step_size": STEP})
    (ref * W.cpu()[:, sel]).sum().backward()
    assert relerr(h_s[:, :n_oracle], ref) < 1e-5
    assert relerr(gy0_s[:n_oracle], z.grad) < 2e-5
    # gradient additivity over contiguous shards
    total = torch.zeros_like(gp)
    for k in range(n_shards):
        lo, hi = (B * k) // n_shards, (B * (k + 1)) // n_shards
        total += solve(everyone[lo:hi])[2]
    ml = slice(13, None)  # ml_net weights and biases (expert_grads=False leaves the 13 scalars at zero)
    assert relerr(total[ml], gp[ml]) < 5e-4
    assert float(gp[:13].abs().max()) == 0.0


def test_properties_small_size_on_host_emulation():
    subprocess.run(["make", "-C", HS_DIR], check=True, capture_output=True)
    lib = L.HodeLib(os.path.join(HS_DIR, "libhode_hostsim.so"),
                    required=["hode_abi_version", "hode_last_error", "hode_fixed_fwd", "hode_fixed_bwd"])
    run_properties(lib, torch.device("cpu"), B=96, n_subset=24, n_shards=4, n_oracle=8)


@pytest.mark.gpu
def test_properties_at_the_full_bench_size():
    run_properties(L.get_lib(), torch.device("cuda:0"), B=1 << 20, n_subset=4096, n_shards=8, n_oracle=48)
