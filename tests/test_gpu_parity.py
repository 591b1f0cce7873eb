"""GPU parity: the CUDA path (through the public Python API, i.e. through the C ABI) against the CPU oracle.

Tolerances (BASELINE.md section 4 / SURVEY.md fact 5), all norm-wise ||a-b||_inf / ||b||_inf:
  fixed-step trajectories 1e-5, gradients 1e-5;  dopri5 at rtol 1e-3/atol 1e-4: identical accept/reject counts;
  dopri5 at the reference's 1e-7/1e-8 (rounding-noise regime): trajectories 1e-4, loss 1e-5, gradients 1e-4 and the
  attempt count within 10 %.
"""
import numpy as np
import pytest
import torch

import hybrid_ode_neurips_2021_b200 as H
from oracle import fields as OF
from oracle import odeint as OI

from _util import EXPERT_NAMES, make_cohort, nan_pattern_equal, oracle_roche, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build_pair(D, seed=0, perturb_scalars=True):
    o = oracle_roche(D, seed, perturb_scalars)
    m = H.RocheODE(D, 1, 14, 1, device=DEV)
    m.load_state_dict(o.state_dict())
    return o, m


def run_both(o, m, y0, a, t, W, **kw):
    o.zero_grad(); m.zero_grad()
    o.set_action(a)
    y0c = y0.clone().requires_grad_(True)
    tr = OI.SolveTrace()
    okw = dict(kw); opts = dict(okw.pop("options", None) or {})
    for k in ("controller", "n_groups", "expert_grads", "tape_capacity"):
        opts.pop(k, None)
    opts["trace"] = tr
    if kw.get("method") == "dopri5":
        opts["differentiable_first_step"] = False
    ref = OI.odeint(o, y0c, t, options=opts, **okw)
    (ref * W).sum().backward()
    m.set_action(a.to(DEV))
    y0g = y0.clone().to(DEV).requires_grad_(True)
    out = H.odeint(m, y0g, t.to(DEV), **kw)
    (out * W.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    return ref, y0c.grad, out, y0g.grad, tr


def check_param_grads(o, m, tol):
    for n in EXPERT_NAMES:
        go, gm = getattr(o, n).grad, getattr(m, n).grad
        assert gm is not None
        assert nan_pattern_equal(go, gm), n
        if not torch.isnan(go).any():
            assert abs(gm.item() - go.item()) <= tol * max(1.0, abs(go.item())), (n, go.item(), gm.item())
    if o.ml_dim > 0:
        assert relerr(m.ml_net[0].weight.grad, o.ml_net[0].weight.grad) < tol
        assert relerr(m.ml_net[0].bias.grad, o.ml_net[0].bias.grad) < tol


@pytest.mark.parametrize("D", [4, 6, 8, 12])
@pytest.mark.parametrize("method,opts", [
    ("rk4", {"step_size": 0.0625}),
    ("rk4", {"step_size": 0.125, "perturb": True}),
    ("midpoint", {"step_size": 0.0625, "perturb": True}),
    ("euler", {"step_size": 0.03125}),
    ("rk4", {"step_size": 0.3}),
])
def test_fixed_grid_parity(D, method, opts):
    B = 37
    o, m = build_pair(D)
    y0, a, _, _ = make_cohort(B, D, seed=D)
    t = torch.arange(0, 15.0)
    W = torch.randn(15, B, D, generator=torch.Generator().manual_seed(1))
    ref, gref, out, gout, tr = run_both(o, m, y0, a, t, W, method=method, options=opts)
    assert out.shape == ref.shape
    assert relerr(out, ref) < 1e-5
    assert torch.allclose(out.cpu(), ref.detach(), rtol=1e-4, atol=1e-5)
    assert relerr(gout, gref) < 1e-5
    check_param_grads(o, m, 2e-5)


def test_fixed_grid_default_grid_is_t():
    # no step_size: the grid is the output grid itself (h = 1 day); small doses keep it finite
    D, B = 6, 5
    o, m = build_pair(D, perturb_scalars=False)
    y0, a, _, _ = make_cohort(B, D, seed=3, dose_max=0.5)
    t = torch.arange(0, 15.0) * 0.25
    W = torch.ones(15, B, D)
    ref, gref, out, gout, _ = run_both(o, m, y0, a, t, W, method="rk4")
    assert relerr(out, ref) < 1e-5 and relerr(gout, gref) < 1e-5


@pytest.mark.parametrize("D", [4, 6, 8, 12])
def test_dopri5_loose_tolerance_identical_step_sequence(D):
    B = 10
    o, m = build_pair(D)
    y0, a, _, _ = make_cohort(B, D, seed=10 + D)
    t = torch.arange(0, 15.0)
    W = torch.randn(15, B, D, generator=torch.Generator().manual_seed(2))
    ref, gref, out, gout, tr = run_both(o, m, y0, a, t, W, method="dopri5", rtol=1e-3, atol=1e-4)
    info = H.last_solve_info()
    assert int(info.accepted[0]) == tr.accepted and int(info.rejected[0]) == tr.rejected
    assert int(info.nfe[0]) == tr.nfe
    assert relerr(out, ref) < 2e-5
    assert relerr(gout, gref) < 5e-5
    check_param_grads(o, m, 1e-4)


@pytest.mark.parametrize("D,B", [(6, 50), (8, 100), (12, 10)])
def test_dopri5_reference_tolerance(D, B):
    o, m = build_pair(D)
    y0, a, _, _ = make_cohort(B, D, seed=20 + D)
    t = torch.arange(0, 15.0)
    W = torch.randn(15, B, D, generator=torch.Generator().manual_seed(3))
    ref, gref, out, gout, tr = run_both(o, m, y0, a, t, W, method="dopri5", rtol=1e-7, atol=1e-8)
    info = H.last_solve_info()
    n_ref = tr.accepted + tr.rejected
    n_out = int(info.accepted[0] + info.rejected[0])
    assert abs(n_out - n_ref) <= 0.10 * n_ref, (n_out, n_ref)
    assert relerr(out, ref) < 1e-4
    lo, lr = (out.cpu() * W).sum().item(), (ref.detach() * W).sum().item()
    assert abs(lo - lr) <= 1e-4 * max(1.0, abs(lr))
    assert relerr(gout, gref) < 1e-4
    if D > 4:
        assert relerr(m.ml_net[0].weight.grad, o.ml_net[0].weight.grad) < 1e-4


def test_dopri5_per_trajectory_equals_batch_one_calls():
    D, B = 8, 6
    o, m = build_pair(D)
    y0, a, _, _ = make_cohort(B, D, seed=5)
    t = torch.arange(0, 15.0)
    m.set_action(a.to(DEV))
    out = H.odeint(m, y0.to(DEV), t.to(DEV), rtol=1e-3, atol=1e-4, method="dopri5", options={"controller": "trajectory"})
    info = H.last_solve_info()
    for b in range(B):
        o.set_action(a[:, b:b + 1])
        tr = OI.SolveTrace()
        with torch.no_grad():
            ref = OI.odeint(o, y0[b:b + 1], t, rtol=1e-3, atol=1e-4, method="dopri5", options={"trace": tr})
        assert int(info.accepted[b]) == tr.accepted and int(info.rejected[b]) == tr.rejected, b
        assert relerr(out[:, b], ref[:, 0]) < 2e-5


def test_dopri5_groups_are_independent_calls():
    D, B, G = 6, 7, 3
    o, m = build_pair(D)
    y0, a, _, _ = make_cohort(B * G, D, seed=6)
    t = torch.arange(0, 15.0)
    m.set_action(a.to(DEV))
    out = H.odeint(m, y0.to(DEV), t.to(DEV), rtol=1e-3, atol=1e-4, method="dopri5", options={"n_groups": G})
    info = H.last_solve_info()
    for g in range(G):
        sl = slice(g * B, (g + 1) * B)
        o.set_action(a[:, sl])
        tr = OI.SolveTrace()
        with torch.no_grad():
            ref = OI.odeint(o, y0[sl], t, rtol=1e-3, atol=1e-4, method="dopri5", options={"trace": tr})
        assert int(info.accepted[g]) == tr.accepted and int(info.rejected[g]) == tr.rejected
        assert relerr(out[:, sl], ref) < 2e-5


def test_multi_warp_group_matches_oracle():
    D, B = 6, 200  # 7 warps in one CTA: exercises the shared-memory stage of the group reduction
    o, m = build_pair(D)
    y0, a, _, _ = make_cohort(B, D, seed=8)
    t = torch.arange(0, 15.0)
    W = torch.ones(15, B, D)
    ref, gref, out, gout, tr = run_both(o, m, y0, a, t, W, method="dopri5", rtol=1e-3, atol=1e-4)
    info = H.last_solve_info()
    assert int(info.accepted[0]) == tr.accepted and int(info.rejected[0]) == tr.rejected
    assert relerr(out, ref) < 2e-5 and relerr(gout, gref) < 5e-5


def test_two_doses_per_patient():
    D, B = 6, 9
    o, m = build_pair(D)
    y0, a, _, _ = make_cohort(B, D, seed=9, n_dose=2)
    t = torch.arange(0, 15.0)
    W = torch.ones(15, B, D)
    ref, gref, out, gout, _ = run_both(o, m, y0, a, t, W, method="rk4", options={"step_size": 0.0625})
    assert tuple(m.times.shape) == (B, 2) and torch.equal(m.times.cpu(), o.times)
    assert relerr(out, ref) < 1e-5 and relerr(gout, gref) < 1e-5
    check_param_grads(o, m, 2e-5)


def test_set_action_matches_reference_loop_and_rejects_ragged():
    T, B = 15, 1000
    _, a, _, _ = make_cohort(B, 6, seed=11)
    a_strided = a.permute(1, 2, 0).contiguous().permute(2, 0, 1)  # the generator's storage order [N][1][T]
    assert not a_strided.is_contiguous()
    dosage, times = OF.dose_schedule(a_strided, 1)
    m = H.RocheODE(6, 1, 14, 1, device=DEV)
    m.set_action(a_strided.to(DEV))
    assert m.times.dtype == torch.int64
    assert torch.equal(m.times.cpu(), times) and torch.equal(m.dosage.cpu(), dosage)
    a2 = a.clone(); a2[:, 0, 0] = 0.0
    with pytest.raises(RuntimeError):
        m.set_action(a2.to(DEV))


@pytest.mark.parametrize("D,obs", [(6, 20), (8, 40), (12, 80), (4, 24)])
@pytest.mark.parametrize("strided", [False, True])
def test_decode_sse_parity(D, obs, strided):
    T, B = 15, 301
    g = torch.Generator().manual_seed(D * obs)
    h = torch.randn(T, B, D, generator=g)
    x = torch.randn(T, B, obs, generator=g)
    mask = (torch.rand(T, B, obs, generator=g) < 0.5).float()
    if strided:  # the reference generator's layout: storage [B][obs][T]
        x = x.permute(1, 2, 0).contiguous().permute(2, 0, 1)
        mask = mask.permute(1, 2, 0).contiguous().permute(2, 0, 1)
    lin = torch.nn.Linear(D, obs)
    hc = h.clone().requires_grad_(True)
    ref = OF.masked_sse(x, lin(hc), mask)
    ref.backward()
    lg = torch.nn.Linear(D, obs).to(DEV); lg.load_state_dict(lin.state_dict())
    hg = h.clone().to(DEV).requires_grad_(True)
    out = H.decode_sse_loss(hg, lg.weight, lg.bias, x.to(DEV), mask.to(DEV))
    out.backward()
    assert abs(out.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert relerr(hg.grad, hc.grad) < 1e-5
    assert relerr(lg.weight.grad, lin.weight.grad) < 2e-5
    assert relerr(lg.bias.grad, lin.bias.grad) < 2e-5


def test_decoder_drop_in_end_to_end():
    D, obs, B = 6, 20, 50
    torch.manual_seed(0)
    od = OF.OracleDecoder(obs, D, rtol=1e-3, atol=1e-4)
    dec = H.RocheExpertDecoder(obs, D, 1, 14, 1, device=DEV)
    assert list(dec.state_dict().keys()) == list(od.state_dict().keys())
    dec.load_state_dict(od.state_dict())
    dec.options["rtol"], dec.options["atol"] = 1e-3, 1e-4
    assert dec.model_name == "HybridDecoder" and torch.equal(dec.t.cpu(), od.t)
    y0, a, x, mask = make_cohort(B, D, obs=obs, seed=12)
    z = y0.clone().requires_grad_(True)
    xh_ref, h_ref = od(z, a)
    loss_ref = OF.masked_sse(x, xh_ref, mask)
    loss_ref.backward()
    zg = y0.clone().to(DEV).requires_grad_(True)
    xh, h = dec(zg, a.to(DEV))
    assert xh.shape == (15, B, obs) and h.shape == (15, B, D)
    loss = H.masked_sse(dec, h, x.to(DEV), mask.to(DEV))
    loss_plain = torch.sum((x.to(DEV) - xh) ** 2 * mask.to(DEV)) / B
    assert abs(loss.item() - loss_plain.item()) <= 1e-5 * abs(loss_plain.item())
    loss.backward()
    assert relerr(xh, xh_ref) < 5e-5
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item())
    assert relerr(zg.grad, z.grad) < 2e-4
    assert relerr(dec.output_function[0].weight.grad, od.output_function[0].weight.grad) < 2e-4
    assert relerr(dec.ode.ml_net[0].weight.grad, od.ode.ml_net[0].weight.grad) < 2e-4


def test_failures_raise_like_torchdiffeq():
    D, B = 6, 4
    o, m = build_pair(D)
    y0, a, _, _ = make_cohort(B, D, seed=13)
    m.set_action(a.to(DEV))
    t = torch.arange(0, 15.0).to(DEV)
    bad = y0.clone(); bad[1, 2] = float("nan")
    with pytest.raises(AssertionError):
        H.odeint(m, bad.to(DEV), t, method="dopri5")
    with pytest.raises(AssertionError, match="max_num_steps"):
        H.odeint(m, y0.to(DEV), t, method="dopri5", options={"max_num_steps": 3})
    with pytest.raises(TypeError):
        H.odeint(torch.nn.Linear(D, D).to(DEV), y0.to(DEV), t)
    with pytest.raises(RuntimeError):
        H.odeint(m, y0, t.cpu())
    with pytest.warns(UserWarning, match="Unexpected arguments"):
        H.odeint(m, y0.to(DEV), t, method="midpoint", options={"step_size": 0.25, "step_t": [1.0]})
