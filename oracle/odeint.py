"""CPU oracle: restatement of ``torchdiffeq==0.2.2``'s ``odeint`` in plain PyTorch.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.

Why a restatement: the reference calls ``from torchdiffeq import odeint as dto`` (``/root/reference/model.py:10``;
call sites ``model.py:837``, ``model.py:842``, ``model.py:1116``) and pins ``torchdiffeq==0.2.2``
(``/root/reference/requirements.txt:9``).  That package is not vendored under ``/root/reference``, is not installed in
this image, is not in ``/opt/wheelhouse`` and cannot be fetched (no network).  So its published algorithm is restated
here from the package's documented behaviour (module names below are the package's own):

* ``_impl/odeint.py``      -> :func:`odeint` (input checks, solver table, time reversal)
* ``_impl/misc.py``        -> :func:`_rms_norm`, :func:`_select_initial_step`, :func:`_compute_error_ratio`,
                              :func:`_optimal_step_size`, :class:`_TimeCast` (``_PerturbFunc``)
* ``_impl/rk_common.py``   -> :func:`_rk_adaptive_attempt`, :class:`AdaptiveRK`
* ``_impl/dopri5.py``      -> :data:`DOPRI5`
* ``_impl/fixed_grid.py``  -> ``euler`` / ``midpoint`` / ``rk4`` (3/8 rule) step functions
* ``_impl/solvers.py``     -> :class:`FixedGrid` (grid construction, linear interpolation to output times)
* ``_impl/interp.py``      -> :func:`_quartic_fit`, :func:`_quartic_eval`
* ``_impl/adjoint.py``     -> :func:`odeint_adjoint` (``OdeintAdjointMethod``, default mixed adjoint norm and ``seminorm``)

PARITY UNPINNED BY THE REFERENCE: the reference repository has no tests and no golden vectors at the solver boundary
(SURVEY.md section 4).  The restatement is pinned instead by (i) closed-form ODEs and order-of-convergence checks,
(ii) the Butcher order conditions of the tableaux, (iii) scipy's independent ``RK45`` (same Dormand-Prince pair), and
(iv) the solver-level values of SURVEY.md Appendix C.2 (``tests/test_oracle_*.py``).  One behaviour could not be
checked against the package source: whether the automatically selected first step carries gradient.  It is
switchable (``options['differentiable_first_step']``, default ``True`` = what the un-decorated
``_select_initial_step`` of the package does) so that its effect can be measured.

Every arithmetic choice that affects results is kept: time-like scalars of the adaptive solver are float64, the
state is ``y0.dtype``; stage times are cast to the state dtype before the vector field sees them; ``alpha == 1``
stages are evaluated one ulp before ``t1``; fixed-grid solvers do all time arithmetic in ``t.dtype``.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import torch

__all__ = ["odeint", "odeint_adjoint", "SolveTrace", "DOPRI5", "SOLVERS"]

NEXT, PREV, NONE = 1, -1, 0


# --------------------------------------------------------------------------------------------------------------
# instrumentation
# --------------------------------------------------------------------------------------------------------------
@dataclass
class SolveTrace:
    """Filled in by :func:`odeint` when passed as ``options['trace']`` (oracle-only option)."""

    nfe: int = 0
    accepted: int = 0
    rejected: int = 0
    first_step: float = float("nan")
    # one entry per attempt: (t0, dt, error_ratio, accepted)
    attempts: List[tuple] = field(default_factory=list)
    # fixed-grid: number of grid steps
    steps: int = 0


# --------------------------------------------------------------------------------------------------------------
# misc.py
# --------------------------------------------------------------------------------------------------------------
def _rms_norm(x: torch.Tensor) -> torch.Tensor:
    # root-mean-square over ALL elements of the state tensor: one scalar per odeint call (batch-coupled control)
    return x.pow(2).mean().sqrt()


class _StraightThrough(torch.autograd.Function):
    """value of ``out``, gradient of ``x`` (the package's ``_StitchGradient``)."""

    @staticmethod
    def forward(ctx, x, out):
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None


def _nextafter(x: torch.Tensor, toward: torch.Tensor) -> torch.Tensor:
    with torch.no_grad():
        out = torch.nextafter(x, toward)
    return _StraightThrough.apply(x, out)


class _TimeCast:
    """The package wraps ``func`` so that ``t`` is cast to the state dtype and optionally moved by one ulp."""

    def __init__(self, func: Callable, trace: Optional[SolveTrace]):
        self.func = func
        self.trace = trace

    def __call__(self, t: torch.Tensor, y: torch.Tensor, perturb: int = NONE) -> torch.Tensor:
        t = t.to(y.dtype)
        if perturb == NEXT:
            t = _nextafter(t, t + 1)
        elif perturb == PREV:
            t = _nextafter(t, t - 1)
        if self.trace is not None:
            self.trace.nfe += 1
        return self.func(t, y)


def _select_initial_step(func, t0, y0, order, rtol, atol, norm, f0):
    """Hairer, Norsett & Wanner, Solving ODEs I, II.4 'starting step size'. All arithmetic in ``y0.dtype``."""
    dtype, device, t_dtype = y0.dtype, y0.device, t0.dtype
    t0 = t0.to(dtype)
    scale = atol + torch.abs(y0) * rtol
    d0 = norm(y0 / scale)
    d1 = norm(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=dtype, device=device)
    else:
        h0 = 0.01 * d0 / d1
    y1 = y0 + h0 * f0
    f1 = func(t0 + h0, y1)
    d2 = norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=dtype, device=device), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    return torch.min(100 * h0, h1).to(t_dtype)


def _compute_error_ratio(err, rtol, atol, y0, y1, norm):
    tol = atol + rtol * torch.max(y0.abs(), y1.abs())
    return norm(err / tol).abs()


@torch.no_grad()
def _optimal_step_size(last_step, error_ratio, safety, ifactor, dfactor, order):
    if error_ratio == 0:
        return last_step * ifactor
    if error_ratio < 1:
        dfactor = torch.ones((), dtype=last_step.dtype, device=last_step.device)
    error_ratio = error_ratio.type_as(last_step)
    exponent = torch.tensor(order, dtype=last_step.dtype, device=last_step.device).reciprocal()
    factor = torch.min(ifactor, torch.max(safety / error_ratio ** exponent, dfactor))
    return last_step * factor


# --------------------------------------------------------------------------------------------------------------
# interp.py
# --------------------------------------------------------------------------------------------------------------
def _quartic_fit(y0, y1, y_mid, f0, f1, dt):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    e = y0
    return [e, d, c, b, a]


def _quartic_eval(coeff, t0, t1, t):
    assert (t0 <= t) & (t <= t1), "invalid interpolation, fails `t0 <= t <= t1`: {}, {}, {}".format(t0, t, t1)
    x = ((t - t0) / (t1 - t0)).to(coeff[0].dtype)
    total = coeff[0] + x * coeff[1]
    x_power = x
    for c in coeff[2:]:
        x_power = x_power * x
        total = total + x_power * c
    return total


# --------------------------------------------------------------------------------------------------------------
# dopri5.py  (Dormand-Prince 5(4), Shampine's dense-output mid-point weights)
# --------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Tableau:
    alpha: tuple
    beta: tuple  # tuple of tuples
    c_sol: tuple
    c_error: tuple
    c_mid: tuple
    order: int


DOPRI5 = Tableau(
    alpha=(1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0),
    beta=(
        (1 / 5,),
        (3 / 40, 9 / 40),
        (44 / 45, -56 / 15, 32 / 9),
        (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
        (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
        (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84),
    ),
    c_sol=(35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0),
    c_error=(
        35 / 384 - 1951 / 21600,
        0.0,
        500 / 1113 - 22642 / 50085,
        125 / 192 - 451 / 720,
        -2187 / 6784 - -12231 / 42400,
        11 / 84 - 649 / 6300,
        -1.0 / 60.0,
    ),
    c_mid=(
        6025192743 / 30085553152 / 2,
        0.0,
        51252292925 / 65400821598 / 2,
        -2691868925 / 45128329728 / 2,
        187940372067 / 1594534317056 / 2,
        -1776094331 / 19743644256 / 2,
        11237099 / 235043384 / 2,
    ),
    order=5,
)


# --------------------------------------------------------------------------------------------------------------
# rk_common.py
# --------------------------------------------------------------------------------------------------------------
def _rk_adaptive_attempt(func, y0, f0, t0, dt, t1, alpha, beta, c_error):
    """One explicit RK attempt. ``alpha/beta/c_error`` are tensors of ``y0.dtype``; ``t0, dt, t1`` are float64."""
    t0 = t0.to(y0.dtype)
    dt = dt.to(y0.dtype)
    t1 = t1.to(y0.dtype)
    ks = [f0]
    yi = y0
    for alpha_i, beta_i in zip(alpha, beta):
        if alpha_i == 1.0:
            ti, perturb = t1, PREV  # "always step to just before the end time, in case of discontinuities"
        else:
            ti, perturb = t0 + alpha_i * dt, NONE
        k = torch.stack(ks, dim=-1)
        yi = y0 + k.matmul(beta_i * dt).view_as(f0)
        ks.append(func(ti, yi, perturb=perturb))
    k = torch.stack(ks, dim=-1)
    # FSAL pair: c_sol[:-1] == beta[-1] and c_sol[-1] == 0, so y1 is the last stage's input and f1 its output
    y1 = yi
    f1 = ks[-1]
    y1_error = k.matmul(dt * c_error)
    return y1, f1, y1_error, k


class AdaptiveRK:
    def __init__(
        self,
        func,
        y0,
        rtol,
        atol,
        tableau: Tableau = DOPRI5,
        first_step=None,
        step_t=None,
        jump_t=None,
        safety=0.9,
        ifactor=10.0,
        dfactor=0.2,
        max_num_steps=2 ** 31 - 1,
        dtype=torch.float64,
        norm=_rms_norm,
        differentiable_first_step=True,
        trace: Optional[SolveTrace] = None,
        **unused,
    ):
        if unused:
            warnings.warn("{}: Unexpected arguments {}".format("Dopri5Solver", sorted(unused)))
        dev = y0.device
        self.func, self.y0, self.norm, self.trace = func, y0, norm, trace
        self.dtype = dtype
        self.rtol = torch.as_tensor(rtol, dtype=dtype, device=dev)
        self.atol = torch.as_tensor(atol, dtype=dtype, device=dev)
        self.first_step = None if first_step is None else torch.as_tensor(first_step, dtype=dtype, device=dev)
        self.safety = torch.as_tensor(safety, dtype=dtype, device=dev)
        self.ifactor = torch.as_tensor(ifactor, dtype=dtype, device=dev)
        self.dfactor = torch.as_tensor(dfactor, dtype=dtype, device=dev)
        self.max_num_steps = max_num_steps
        self.step_t = None if step_t is None else torch.as_tensor(step_t, dtype=dtype, device=dev)
        self.jump_t = None if jump_t is None else torch.as_tensor(jump_t, dtype=dtype, device=dev)
        self.differentiable_first_step = differentiable_first_step
        self.order = tableau.order
        sd = y0.dtype
        # tableau tensors are built in float64 and cast to the state dtype
        self.alpha = torch.tensor(tableau.alpha, dtype=torch.float64).to(device=dev, dtype=sd)
        self.beta = [torch.tensor(b, dtype=torch.float64).to(device=dev, dtype=sd) for b in tableau.beta]
        self.c_error = torch.tensor(tableau.c_error, dtype=torch.float64).to(device=dev, dtype=sd)
        self.c_mid = torch.tensor(tableau.c_mid, dtype=torch.float64).to(device=dev, dtype=sd)

    # -- setup ------------------------------------------------------------------------------------------------
    def _before_integrate(self, t):
        t0 = t[0]
        f0 = self.func(t0, self.y0)
        if self.first_step is None:
            if self.differentiable_first_step:
                first = _select_initial_step(
                    self.func, t0, self.y0, self.order - 1, self.rtol, self.atol, self.norm, f0
                )
            else:
                with torch.no_grad():
                    first = _select_initial_step(
                        self.func, t0, self.y0.detach(), self.order - 1, self.rtol, self.atol, self.norm, f0.detach()
                    )
        else:
            first = self.first_step
        if self.trace is not None:
            self.trace.first_step = float(first.detach())
        # (y1, f1, t0, t1, dt, interp_coeff)
        self.state = (self.y0, f0, t0, t0, first, [self.y0] * 5)

        def _prep(pts):
            if pts is None:
                return torch.tensor([], dtype=self.dtype, device=self.y0.device)
            pts = torch.sort(pts.flatten()).values
            return pts[pts > t0]

        self.step_pts = _prep(self.step_t)
        self.jump_pts = _prep(self.jump_t)
        self.next_step_index = 0
        self.next_jump_index = 0

    # -- one attempt ------------------------------------------------------------------------------------------
    def _adaptive_step(self, state):
        y0, f0, _, t0, dt, coeff = state
        t1 = t0 + dt
        assert t0 + dt > t0, "underflow in dt {}".format(dt.item())
        assert torch.isfinite(y0).all(), "non-finite values in state `y`: {}".format(y0)

        on_step_t = False
        if len(self.step_pts):
            nxt = self.step_pts[self.next_step_index]
            on_step_t = bool(t0 < nxt < t0 + dt)
            if on_step_t:
                t1 = nxt
                dt = t1 - t0
        on_jump_t = False
        if len(self.jump_pts):
            nxt = self.jump_pts[self.next_jump_index]
            on_jump_t = bool(t0 < nxt < t0 + dt)
            if on_jump_t:
                on_step_t = False
                t1 = nxt
                dt = t1 - t0

        y1, f1, y1_error, k = _rk_adaptive_attempt(self.func, y0, f0, t0, dt, t1, self.alpha, self.beta, self.c_error)
        ratio = _compute_error_ratio(y1_error, self.rtol, self.atol, y0, y1, self.norm)
        accept = bool(ratio <= 1)
        if self.trace is not None:
            self.trace.attempts.append((float(t0.detach() if isinstance(t0, torch.Tensor) else t0), float(dt.detach()), float(ratio.detach()), accept))
            if accept:
                self.trace.accepted += 1
            else:
                self.trace.rejected += 1

        if accept:
            t_next, y_next = t1, y1
            dts = dt.type_as(y0)
            y_mid = y0 + k.matmul(dts * self.c_mid).view_as(y0)
            coeff = _quartic_fit(y0, y1, y_mid, k[..., 0], k[..., -1], dts)
            if on_step_t and self.next_step_index != len(self.step_pts) - 1:
                self.next_step_index += 1
            if on_jump_t:
                if self.next_jump_index != len(self.jump_pts) - 1:
                    self.next_jump_index += 1
                f1 = self.func(t_next, y_next, perturb=NEXT)
            f_next = f1
        else:
            t_next, y_next, f_next = t0, y0, f0
        dt_next = _optimal_step_size(dt, ratio, self.safety, self.ifactor, self.dfactor, self.order)
        return (y_next, f_next, t0, t_next, dt_next, coeff)

    # -- driver -----------------------------------------------------------------------------------------------
    def integrate(self, t):
        sol = [self.y0]
        t = t.to(self.dtype)
        self._before_integrate(t)
        for i in range(1, len(t)):
            n_steps = 0
            while t[i] > self.state[3]:
                assert n_steps < self.max_num_steps, "max_num_steps exceeded ({}>={})".format(
                    n_steps, self.max_num_steps
                )
                self.state = self._adaptive_step(self.state)
                n_steps += 1
            sol.append(_quartic_eval(self.state[5], self.state[2], self.state[3], t[i]))
        return torch.stack(sol, dim=0)


# --------------------------------------------------------------------------------------------------------------
# fixed_grid.py / solvers.py
# --------------------------------------------------------------------------------------------------------------
_ONE_THIRD = 1 / 3
_TWO_THIRDS = 2 / 3


def _euler_step(func, t0, dt, t1, y0, perturb):
    f0 = func(t0, y0, perturb=NEXT if perturb else NONE)
    return dt * f0


def _midpoint_step(func, t0, dt, t1, y0, perturb):
    half_dt = 0.5 * dt
    f0 = func(t0, y0, perturb=NEXT if perturb else NONE)
    y_mid = y0 + f0 * half_dt
    return dt * func(t0 + half_dt, y_mid)


def _rk4_38_step(func, t0, dt, t1, y0, perturb):
    # torchdiffeq's `rk4` is the 3/8 rule ("smaller error with slightly more compute"), not the classical scheme
    k1 = func(t0, y0, perturb=NEXT if perturb else NONE)
    k2 = func(t0 + dt * _ONE_THIRD, y0 + dt * k1 * _ONE_THIRD)
    k3 = func(t0 + dt * _TWO_THIRDS, y0 + dt * (k2 - k1 * _ONE_THIRD))
    k4 = func(t1, y0 + dt * (k1 - k2 + k3), perturb=PREV if perturb else NONE)
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


_FIXED_STEPS = {"euler": _euler_step, "midpoint": _midpoint_step, "rk4": _rk4_38_step}
_FIXED_NAMES = {"euler": "Euler", "midpoint": "Midpoint", "rk4": "RK4"}


def fixed_grid_points(t: torch.Tensor, step_size) -> torch.Tensor:
    """Grid of a fixed-step solve: ``t`` itself, or ``arange(niters) * step_size + t[0]`` with the end clamped."""
    if step_size is None:
        return t
    start, end = t[0], t[-1]
    niters = torch.ceil((end - start) / step_size + 1).item()
    grid = torch.arange(0, niters, dtype=t.dtype, device=t.device) * step_size + start
    grid[-1] = t[-1]
    return grid


class FixedGrid:
    def __init__(
        self,
        func,
        y0,
        method,
        step_size=None,
        grid_constructor=None,
        interp="linear",
        perturb=False,
        trace: Optional[SolveTrace] = None,
        rtol=None,
        atol=None,
        norm=None,
        differentiable_first_step=None,
        **unused,
    ):
        if unused:
            warnings.warn("{}: Unexpected arguments {}".format(_FIXED_NAMES[method], sorted(unused)))
        if step_size is not None and grid_constructor is not None:
            raise ValueError("step_size and grid_constructor are mutually exclusive arguments.")
        if interp != "linear":
            raise ValueError("Unknown interpolation method {}".format(interp))
        self.func, self.y0, self.step = func, y0, _FIXED_STEPS[method]
        self.step_size, self.grid_constructor, self.perturb, self.trace = step_size, grid_constructor, perturb, trace

    def integrate(self, t):
        if self.grid_constructor is not None:
            grid = self.grid_constructor(self.func, self.y0, t)
        else:
            grid = fixed_grid_points(t, self.step_size)
        assert grid[0] == t[0] and grid[-1] == t[-1]
        sol = [self.y0]
        j = 1
        y0 = self.y0
        for t0, t1 in zip(grid[:-1], grid[1:]):
            dt = t1 - t0
            y1 = y0 + self.step(self.func, t0, dt, t1, y0, self.perturb)
            if self.trace is not None:
                self.trace.steps += 1
            while j < len(t) and t1 >= t[j]:
                if t[j] == t0:
                    sol.append(y0)
                elif t[j] == t1:
                    sol.append(y1)
                else:
                    slope = (t[j] - t0) / (t1 - t0)
                    sol.append(y0 + slope * (y1 - y0))
                j += 1
            y0 = y1
        return torch.stack(sol, dim=0)


# --------------------------------------------------------------------------------------------------------------
# odeint.py
# --------------------------------------------------------------------------------------------------------------
SOLVERS = ("dopri5", "euler", "midpoint", "rk4")


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    """``torchdiffeq.odeint`` for a tensor state.  Returns ``[len(t), *y0.shape]``.

    Only the solvers the reference's experiments select are restated: ``dopri5`` (``sim_config.py:50``),
    ``midpoint`` / ``rk4`` (``experiments/real.sh:9-17``) and ``euler``.
    """
    if event_fn is not None:
        raise NotImplementedError("event handling is not used by the reference and not restated")
    if not isinstance(y0, torch.Tensor):
        raise NotImplementedError("tuple states are not used by the reference and not restated")
    assert isinstance(t, torch.Tensor) and t.ndimension() == 1, "t must be one dimensional"
    assert torch.is_floating_point(t), "t must be a floating point Tensor"
    assert torch.is_floating_point(y0), "`y0` must be a floating point Tensor"
    assert not t.requires_grad or True
    for name, tol in (("rtol", rtol), ("atol", atol)):
        if isinstance(tol, torch.Tensor):
            assert not tol.requires_grad, name + " cannot require gradient"
    increasing = bool((t[1:] > t[:-1]).all())
    decreasing = bool((t[1:] < t[:-1]).all())
    assert increasing or decreasing, "t must be strictly increasing or decreasing"
    options = {} if options is None else dict(options)
    trace = options.pop("trace", None)
    if method is None:
        method = "dopri5"
    if method not in SOLVERS:
        raise ValueError('Invalid method "{}". Must be one of {}'.format(method, "{" + ", ".join(SOLVERS) + "}"))
    t = t.to(y0.device)

    user_func = func
    if decreasing and not increasing:
        t = -t
        user_func = lambda tt, yy: -func(-tt, yy)  # noqa: E731
    wrapped = _TimeCast(user_func, trace)

    if method == "dopri5":
        solver = AdaptiveRK(wrapped, y0, rtol, atol, trace=trace, **options)
    else:
        solver = FixedGrid(wrapped, y0, method, trace=trace, **options)
    return solver.integrate(t)


# --------------------------------------------------------------------------------------------------------------
# adjoint.py
# --------------------------------------------------------------------------------------------------------------
class _AdjointMethod(torch.autograd.Function):
    """``OdeintAdjointMethod``: forward = plain ``odeint`` under ``no_grad``; backward integrates the augmented system
    ``(vjp_t, y, adj_y, *adj_params)`` backwards over every output interval with ``odeint`` itself (decreasing ``t`` ->
    the time-negation path of :func:`odeint`), resetting ``y`` to the stored forward solution at every output time."""

    @staticmethod
    def forward(ctx, func, y0, t, rtol, atol, method, options, adj_rtol, adj_atol, adj_method, adj_options, n_params,
                *adjoint_params):
        ctx.func, ctx.cfg = func, (adj_rtol, adj_atol, adj_method, adj_options)
        with torch.no_grad():
            y = odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
        ctx.save_for_backward(t, y, *adjoint_params)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        func = ctx.func
        adj_rtol, adj_atol, adj_method, adj_options = ctx.cfg
        t, y, *adjoint_params = ctx.saved_tensors
        adjoint_params = tuple(adjoint_params)
        with torch.no_grad():
            # the package keeps a tuple state and flattens it (``_TupleFunc``) before the solver sees it
            state = [torch.zeros((), dtype=y.dtype, device=y.device), y[-1], grad_y[-1]]
            state.extend(torch.zeros_like(p) for p in adjoint_params)
            shapes = [s.shape for s in state]
            sizes = [s.numel() for s in state]

            def flatten(parts):
                return torch.cat([p.reshape(-1) for p in parts])

            def unflatten(flat):
                out, off = [], 0
                for shp, n in zip(shapes, sizes):
                    out.append(flat[off:off + n].view(shp))
                    off += n
                return out

            def augmented_dynamics(tt, flat):
                parts = unflatten(flat)
                yy, adj_y = parts[1], parts[2]
                with torch.enable_grad():
                    t_ = tt.detach()
                    yy = yy.detach().requires_grad_(True)
                    func_eval = func(t_, yy)
                    vjps = torch.autograd.grad(func_eval, (yy,) + adjoint_params, -adj_y, allow_unused=True,
                                               retain_graph=True)
                vjp_y = torch.zeros_like(yy) if vjps[0] is None else vjps[0]
                vjp_params = [torch.zeros_like(p) if v is None else v for p, v in zip(adjoint_params, vjps[1:])]
                # vjp_t: t carries no gradient request (t_requires_grad is False at every reference call site)
                return flatten([torch.zeros_like(t_).to(yy.dtype), func_eval.detach(), vjp_y] + vjp_params)

            # adaptive adjoint solves: the package's default adjoint norm (adjoint.py handle_adjoint_norm_) is a MIXED norm
            # over the tuple, max(|vjp_t|, rms(y), rms(adj_y), max_k rms(adj_param_k)); 'seminorm' drops the parameter part
            adj_options = dict(adj_options)
            if adj_method == "dopri5":
                seminorm = adj_options.get("norm") == "seminorm"
                if "norm" not in adj_options or seminorm:
                    def adjoint_norm(flat):
                        parts = unflatten(flat)
                        terms = [parts[0].abs(), _rms_norm(parts[1]), _rms_norm(parts[2])]
                        if not seminorm:
                            terms += [_rms_norm(p) for p in parts[3:] if p.numel() > 0]
                        return torch.stack([x.reshape(()) for x in terms]).max()

                    adj_options["norm"] = adjoint_norm
            for i in range(len(t) - 1, 0, -1):
                sol = odeint(augmented_dynamics, flatten(state), t[i - 1:i + 1].flip(0), rtol=adj_rtol, atol=adj_atol,
                             method=adj_method, options=adj_options)
                state = unflatten(sol[1])
                state[1] = y[i - 1]
                state[2] = state[2] + grad_y[i - 1]
            adj_y, adj_params = state[2], state[3:]
        return (None, adj_y, None, None, None, None, None, None, None, None, None, None, *adj_params)


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None, adjoint_rtol=None,
                   adjoint_atol=None, adjoint_method=None, adjoint_options=None, adjoint_params=None):
    """``torchdiffeq.odeint_adjoint`` (the import the reference keeps commented out at ``model.py:9``).  Fixed-grid methods
    and ``dopri5`` (the adaptive adjoint solve is controlled by the package's mixed norm over the augmented tuple, or by
    ``adjoint_options={'norm': 'seminorm'}``; the CUDA path has kernels for the fixed-grid methods only).  Defaults follow
    the package: the adjoint solve uses the forward method / tolerances / options unless overridden; ``adjoint_params``
    defaults to ``func.parameters()``, filtered to those that require grad."""
    if event_fn is not None:
        raise NotImplementedError("event handling is not used by the reference and not restated")
    if adjoint_params is None and not isinstance(func, torch.nn.Module):
        raise ValueError("func must be an instance of nn.Module to specify the adjoint parameters; alternatively they "
                         "can be specified explicitly via the `adjoint_params` argument.")
    method = "dopri5" if method is None else method
    adjoint_rtol = rtol if adjoint_rtol is None else adjoint_rtol
    adjoint_atol = atol if adjoint_atol is None else adjoint_atol
    adjoint_method = method if adjoint_method is None else adjoint_method
    if adjoint_options is None:
        adjoint_options = {k: v for k, v in options.items() if k != "norm"} if options is not None else {}
    else:
        adjoint_options = dict(adjoint_options)
    adjoint_params = tuple(func.parameters()) if adjoint_params is None else tuple(adjoint_params)
    adjoint_params = tuple(p for p in adjoint_params if p.requires_grad)
    return _AdjointMethod.apply(func, y0, t, rtol, atol, method, options, adjoint_rtol, adjoint_atol, adjoint_method,
                                adjoint_options, len(adjoint_params), *adjoint_params)
