"""The torchdiffeq restatement (oracle/odeint.py) against mathematics and against the solver-level values of
SURVEY.md Appendix C.2.  The reference has no tests of its own at this boundary ("parity unpinned")."""
import math

import numpy as np
import pytest
import torch

from oracle import fields as OF
from oracle import odeint as OI


class Decay(torch.nn.Module):
    def forward(self, t, y):
        return -y


class Oscillator(torch.nn.Module):
    def forward(self, t, y):
        return torch.stack([y[..., 1], -y[..., 0]], dim=-1)


class Poly(torch.nn.Module):
    """y' = 4 t^3 - 3 t^2 + 2 t - 1  ->  y = t^4 - t^3 + t^2 - t + y0"""

    def forward(self, t, y):
        return (4 * t ** 3 - 3 * t ** 2 + 2 * t - 1).expand_as(y).to(y.dtype)


def test_tableau_order_conditions():
    T = OI.DOPRI5
    a = np.array(T.alpha)
    for i, row in enumerate(T.beta):
        assert abs(sum(row) - a[i]) < 1e-15
    b = np.array(T.c_sol)
    c = np.array([0.0] + list(a))
    assert abs(b.sum() - 1) < 1e-15 and abs((b * c).sum() - 0.5) < 1e-15
    assert abs((b * c ** 2).sum() - 1 / 3) < 1e-15 and abs((b * c ** 3).sum() - 1 / 4) < 1e-15
    assert abs((b * c ** 4).sum() - 1 / 5) < 1e-15
    assert abs(sum(T.c_error)) < 1e-15
    m = np.array(T.c_mid)
    assert abs(m.sum() - 0.5) < 1e-15 and abs((m * c).sum() - 1 / 8) < 1e-12 and abs((m * c ** 2).sum() - 1 / 24) < 1e-12


@pytest.mark.parametrize("method,order", [("euler", 1), ("midpoint", 2), ("rk4", 4)])
def test_fixed_grid_order_of_convergence(method, order):
    y0 = torch.tensor([[1.0, 0.0]], dtype=torch.float64)
    t = torch.tensor([0.0, 1.0], dtype=torch.float64)
    errs = []
    for h in (0.1, 0.05):
        y = OI.odeint(Oscillator(), y0, t, method=method, options={"step_size": h})[-1, 0]
        errs.append(float(torch.linalg.norm(y - torch.tensor([math.cos(1.0), -math.sin(1.0)], dtype=torch.float64))))
    rate = math.log2(errs[0] / errs[1])
    assert abs(rate - order) < 0.25, (errs, rate)


def test_rk4_is_the_three_eighths_rule():
    # one step of y' = y with h = 1: both 4th-order rules give 1 + 1 + 1/2 + 1/6 + 1/24, so use a t-dependent field
    class F(torch.nn.Module):
        def forward(self, t, y):
            return y * t

    y0 = torch.tensor([[1.0]], dtype=torch.float64)
    h = 0.5
    got = OI.odeint(F(), y0, torch.tensor([0.0, h], dtype=torch.float64), method="rk4")[-1].item()
    k1 = 0.0
    k2 = (1 + h * k1 / 3) * (h / 3)
    k3 = (1 + h * (k2 - k1 / 3)) * (2 * h / 3)
    k4 = (1 + h * (k1 - k2 + k3)) * h
    assert abs(got - (1 + h * (k1 + 3 * (k2 + k3) + k4) / 8)) < 1e-15


def test_fixed_grid_construction_and_linear_interpolation():
    t = torch.tensor([0.0, 0.3, 1.0])
    g = OI.fixed_grid_points(t, 0.25)
    assert torch.equal(g, torch.tensor([0.0, 0.25, 0.5, 0.75, 1.0]))
    g = OI.fixed_grid_points(t, 0.4)
    assert torch.allclose(g, torch.tensor([0.0, 0.4, 0.8, 1.0])) and g[-1] == 1.0
    tr = OI.SolveTrace()
    y = OI.odeint(Decay(), torch.ones(1, 1), t, method="euler", options={"step_size": 0.25, "trace": tr})
    assert tr.steps == 4 and tr.nfe == 4
    # t = 0.3 lies in [0.25, 0.5]: linear interpolation between the Euler states 0.75 and 0.5625
    assert abs(y[1].item() - (0.75 + (0.3 - 0.25) / 0.25 * (0.5625 - 0.75))) < 1e-6
    assert abs(y[2].item() - 0.75 ** 4) < 1e-6


def test_perturb_moves_first_and_last_stage_times_by_one_ulp():
    seen = []

    class Rec(torch.nn.Module):
        def forward(self, t, y):
            seen.append(float(t))
            return -y

    OI.odeint(Rec(), torch.ones(1, 1), torch.tensor([1.0, 2.0]), method="rk4", options={"perturb": True})
    one, two = np.float32(1.0), np.float32(2.0)
    assert seen[0] == float(np.nextafter(one, np.float32(3.0))) and seen[3] == float(np.nextafter(two, np.float32(0.0)))
    assert seen[1] == float(np.float32(1.0) + np.float32(1.0) * np.float32(1 / 3))


def test_dopri5_accuracy_counts_and_dense_output():
    t = torch.linspace(0, 5, 11, dtype=torch.float64)
    tr = OI.SolveTrace()
    y = OI.odeint(Oscillator(), torch.tensor([[1.0, 0.0]], dtype=torch.float64), t, rtol=1e-8, atol=1e-10,
                  options={"trace": tr})
    exact = torch.stack([torch.cos(t), -torch.sin(t)], dim=-1)[:, None]
    assert float((y - exact).abs().max()) < 1e-7
    assert tr.nfe == 2 + 6 * (tr.accepted + tr.rejected)
    # a quartic solution is reproduced exactly by the 5th-order step and by the quartic dense output
    tq = torch.tensor([0.0, 0.13, 0.5, 0.77, 1.0], dtype=torch.float64)
    yq = OI.odeint(Poly(), torch.zeros(1, 1, dtype=torch.float64), tq, rtol=1e-6, atol=1e-8, options={"first_step": 1.0})
    exact = tq ** 4 - tq ** 3 + tq ** 2 - tq
    assert float((yq[:, 0, 0] - exact).abs().max()) < 1e-12


def test_dopri5_against_scipy_rk45():
    from scipy.integrate import solve_ivp

    sol = solve_ivp(lambda t, y: [y[1], -y[0]], (0, 5), [1.0, 0.0], method="RK45", rtol=1e-9, atol=1e-11, t_eval=[5.0])
    y = OI.odeint(Oscillator(), torch.tensor([[1.0, 0.0]], dtype=torch.float64), torch.tensor([0.0, 5.0], dtype=torch.float64),
                  rtol=1e-9, atol=1e-11)
    assert np.allclose(y[-1, 0].numpy(), sol.y[:, -1], atol=1e-8)


def test_controller_constants():
    one = torch.tensor(1.0, dtype=torch.float64)
    f = lambda r: float(OI._optimal_step_size(one, torch.tensor(r, dtype=torch.float32), torch.tensor(0.9, dtype=torch.float64),
                                              torch.tensor(10.0, dtype=torch.float64), torch.tensor(0.2, dtype=torch.float64), 5))
    assert f(0.0) == 10.0
    assert abs(f(1e-10) - 10.0) < 1e-12              # clipped by ifactor
    assert abs(f(0.5) - 0.9 / 0.5 ** 0.2) < 1e-7
    assert abs(f(0.99) - max(0.9 / 0.99 ** 0.2, 1.0)) < 1e-7   # accepted steps never shrink
    assert abs(f(1e6) - 0.2) < 1e-12                 # clipped by dfactor


def test_unknown_options_warn_and_errors():
    with pytest.warns(UserWarning, match="Unexpected arguments"):
        OI.odeint(Decay(), torch.ones(1, 1), torch.tensor([0.0, 1.0]), method="midpoint", options={"step_t": [0.5]})
    with pytest.raises(ValueError):
        OI.odeint(Decay(), torch.ones(1, 1), torch.tensor([0.0, 1.0]), method="nope")
    with pytest.raises(AssertionError, match="underflow in dt"):
        OI.odeint(Decay(), torch.full((1, 1), float("nan")), torch.tensor([0.0, 1.0]))
    with pytest.raises(AssertionError, match="max_num_steps"):
        OI.odeint(Oscillator(), torch.tensor([[1.0, 0.0]]), torch.tensor([0.0, 50.0]), options={"max_num_steps": 2})


# ---- hybrid RocheODE(6), solver-level values (float64 runs) -----------------------------
def _c2_setup(dtype):
    ode = OF.OracleRocheODE(6)
    with torch.no_grad():
        ode.ml_net[0].weight.copy_(torch.arange(12.0).view(2, 6) / 10 - 0.5)
        ode.ml_net[0].bias.copy_(torch.tensor([0.1, -0.2]))
    ode = ode.to(dtype)
    a = torch.zeros(15, 2, 1, dtype=dtype)
    a[3, 0, 0] = 2
    a[0, 1, 0] = 5
    ode.set_action(a)
    y0 = torch.tensor([[0.02, 0.01, 0.03, 0, 0.05, 0.01], [0.01, 0.02, 0.03, 0.04, 0, 0.02]], dtype=dtype)
    return ode, y0, torch.arange(0, 15, dtype=dtype)


C2 = [
    ("rk4", dict(options={"step_size": 0.0625}), (224, 896),
     [0.55622529, 0.40016873, 0.53304647, 0.74150683, -0.41406194, -1.37618939],
     [0.0, 0.00124767, 3.29294407, 5.824e-05, -5.49260385, 12.30103122]),
    ("midpoint", dict(options={"step_size": 0.0625, "perturb": True}), (224, 448),
     [0.55548137, 0.40161919, 0.53343739, 0.73487404, -0.41225781, -1.37867994],
     [0.0, 0.00125349, 3.29448268, 5.843e-05, -5.49431191, 12.30167931]),
    ("dopri5", dict(rtol=1e-7, atol=1e-8), (105, 49, 926),
     [0.55535701, 0.40194467, 0.53422358, 0.73576141, -0.41283503, -1.3794567],
     [-0.0, 0.00124767, 3.29294517, 5.824e-05, -5.49260564, 12.30103192]),
    ("dopri5", dict(rtol=1e-3, atol=1e-4), (34, 15, 296),
     [0.56294866, 0.38785483, 0.5238759, 0.78541513, -0.42368672, -1.35151702],
     [-0.00010212, 0.00126448, 3.29729554, 5.825e-05, -5.49893438, 12.3035955]),
]


@pytest.mark.parametrize("method,kw,counts,p0_t4,p1_t14", C2)
def test_appendix_c2_solver_values(method, kw, counts, p0_t4, p1_t14):
    ode, y0, t = _c2_setup(torch.float64)
    tr = OI.SolveTrace()
    kw = dict(kw)
    opts = dict(kw.pop("options", {}), trace=tr)
    with torch.no_grad():
        s = OI.odeint(ode, y0, t, method=method, options=opts, **kw)
    if method == "dopri5":
        assert (tr.accepted, tr.rejected, tr.nfe) == counts
    else:
        assert (tr.steps, tr.nfe) == counts
    assert np.allclose(s[4, 0].numpy(), p0_t4, atol=2e-8)
    assert np.allclose(s[14, 1].numpy(), p1_t14, atol=2e-8)


def test_appendix_c2_loose_dopri5_step_sequence():
    ode, y0, t = _c2_setup(torch.float64)
    tr = OI.SolveTrace()
    with torch.no_grad():
        OI.odeint(ode, y0, t, rtol=1e-3, atol=1e-4, options={"trace": tr})
    dts = [a[1] for a in tr.attempts[:6]]
    assert np.allclose(dts, [0.0179834, 0.17983396, 0.60070702, 0.6442699, 0.96499208, 1.33071374], atol=1e-8)
    assert [a[3] for a in tr.attempts[:12]] == [True] * 5 + [False] * 5 + [True, False]


def test_float32_run_keeps_time_in_float64_and_casts_per_stage():
    ode, y0, t = _c2_setup(torch.float32)
    seen = []
    orig = ode.forward
    ode.forward = lambda tt, yy: (seen.append(tt.dtype), orig(tt, yy))[1]
    tr = OI.SolveTrace()
    with torch.no_grad():
        OI.odeint(ode, y0, t, rtol=1e-3, atol=1e-4, options={"trace": tr})
    assert set(seen) == {torch.float32}
    assert (tr.accepted, tr.rejected) == (34, 15)


# ---- odeint_adjoint restatement (torchdiffeq adjoint.py) -------------------------------------------------------------
class _Decay(torch.nn.Module):
    """dy/dt = -theta * y + cos(t) * b: closed-form sensitivities for theta (b = 0) and a time-dependent forcing."""

    def __init__(self, theta, b=0.0):
        super().__init__()
        self.theta = torch.nn.Parameter(torch.tensor(theta, dtype=torch.float64))
        self.b = torch.nn.Parameter(torch.tensor(b, dtype=torch.float64))

    def forward(self, t, y):
        return -self.theta * y + torch.cos(t) * self.b


@pytest.mark.parametrize("method,h,tol", [("rk4", 0.05, 1e-7), ("midpoint", 0.01, 1e-4), ("euler", 0.001, 2e-3)])
def test_adjoint_gradients_match_closed_form(method, h, tol):
    """L = sum_i w_i y(t_i), y = y0 exp(-theta t): dL/dy0 = sum w_i exp(-theta t_i), dL/dtheta = -sum w_i t_i y(t_i)."""
    f = _Decay(0.7)
    y0 = torch.tensor([[1.3], [0.4]], dtype=torch.float64, requires_grad=True)
    t = torch.tensor([0.0, 0.5, 1.25, 2.0], dtype=torch.float64)
    w = torch.tensor([0.3, -1.0, 2.0, 0.5], dtype=torch.float64)
    out = OI.odeint_adjoint(f, y0, t, method=method, options={"step_size": h})
    (out[:, :, 0] * w[:, None]).sum().backward()
    e = torch.exp(-0.7 * t)
    assert torch.allclose(y0.grad[:, 0], (w * e).sum().expand(2), rtol=tol, atol=0)
    expect = -(w[:, None] * t[:, None] * e[:, None] * y0.detach()[None, :, 0]).sum()
    assert abs(f.theta.grad.item() - expect.item()) <= tol * abs(expect.item())
    # b enters through the forcing only: dL/db = sum_i w_i int_0^{t_i} exp(-theta (t_i - s)) cos(s) ds per trajectory
    s = torch.linspace(0, 2, 200001, dtype=torch.float64)
    db = 0.0
    for wi, ti in zip(w.tolist(), t.tolist()):
        m = s <= ti
        db += wi * torch.trapezoid(torch.exp(-0.7 * (ti - s[m])) * torch.cos(s[m]), s[m]).item()
    assert abs(f.b.grad.item() - 2 * db) <= max(tol, 1e-6) * abs(2 * db)


def test_adjoint_equals_discrete_backprop_to_the_order_of_the_method():
    """optimise-then-discretise vs discretise-then-optimise on a smooth nonlinear field: the gap must fall with the
    method's order (rk4: ~h^4) and the forward values must be IDENTICAL (the adjoint's forward pass is plain odeint)."""

    class F(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.lin = torch.nn.Linear(3, 3).double()

        def forward(self, t, y):
            return torch.tanh(self.lin(y)) * torch.cos(t)

    t = torch.linspace(0, 2, 5, dtype=torch.float64)
    w = torch.randn(5, 4, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    gaps = []
    for h in (0.25, 0.125):
        f = F()
        y0 = torch.randn(4, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
        out = OI.odeint(f, y0, t, method="rk4", options={"step_size": h})
        (out * w).sum().backward()
        g_d = [y0.grad.clone(), f.lin.weight.grad.clone()]
        y0.grad = None
        f.zero_grad()
        out_a = OI.odeint_adjoint(f, y0, t, method="rk4", options={"step_size": h})
        assert torch.equal(out_a, out)
        (out_a * w).sum().backward()
        gaps.append(max(((y0.grad - g_d[0]).norm() / g_d[0].norm()).item(),
                        ((f.lin.weight.grad - g_d[1]).norm() / g_d[1].norm()).item()))
    assert gaps[1] < gaps[0] / 8 and gaps[1] < 1e-6  # 4th order: halving h cuts the gap by ~16


def test_adjoint_argument_handling():
    f = _Decay(0.5)
    y0 = torch.ones(1, 1, dtype=torch.float64)
    t = torch.tensor([0.0, 1.0], dtype=torch.float64)
    with pytest.raises(ValueError):
        OI.odeint_adjoint(lambda tt, yy: -yy, y0, t, method="rk4")  # not an nn.Module and no adjoint_params
    # parameters that do not require grad are dropped from the augmented state
    f.b.requires_grad_(False)
    y0g = y0.clone().requires_grad_(True)
    OI.odeint_adjoint(f, y0g, t, method="rk4", options={"step_size": 0.1}).sum().backward()
    assert f.b.grad is None and f.theta.grad is not None


@pytest.mark.parametrize("adjoint_options", [None, {"norm": "seminorm"}])
def test_adaptive_adjoint_matches_discrete_backprop_at_tight_tolerance(adjoint_options):
    """dopri5 adjoint (mixed adjoint norm over the augmented tuple, or the seminorm) against autograd through the dopri5
    forward solve: at rtol 1e-9 both are the exact sensitivities to ~1e-7.  The forward values are those of plain odeint.
    (Oracle-level groundwork: the CUDA path has no adaptive adjoint kernel yet and raises NotImplementedError.)"""

    class F(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.lin = torch.nn.Linear(3, 3).double()

        def forward(self, t, y):
            return torch.tanh(self.lin(y)) * torch.cos(t)

    t = torch.linspace(0, 2, 5, dtype=torch.float64)
    w = torch.randn(5, 4, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    f = F()
    y0 = torch.randn(4, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
    out = OI.odeint(f, y0, t, rtol=1e-9, atol=1e-10, method="dopri5")
    (out * w).sum().backward()
    g_d = [y0.grad.clone(), f.lin.weight.grad.clone(), f.lin.bias.grad.clone()]
    y0.grad = None
    f.zero_grad()
    out_a = OI.odeint_adjoint(f, y0, t, rtol=1e-9, atol=1e-10, method="dopri5", adjoint_options=adjoint_options)
    assert torch.equal(out_a, out)
    (out_a * w).sum().backward()
    for got, ref in zip([y0.grad, f.lin.weight.grad, f.lin.bias.grad], g_d):
        assert ((got - ref).norm() / ref.norm()).item() < 1e-6
