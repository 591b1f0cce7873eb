"""``odeint`` with torchdiffeq's call signature, executed by the fused sm_100a kernels of ``libhode_b200.so``.

Replaces, for the reference's two vector fields, what ``from torchdiffeq import odeint as dto`` provides at
``/root/reference/model.py:10`` (call sites ``model.py:837, 842, 1116``): the whole solve is ONE kernel launch
(forward) plus ONE launch for the reverse sweep, exposed to autograd as a custom ``torch.autograd.Function``.

Semantics kept from torchdiffeq 0.2.2: ``odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None,
event_fn=None) -> Tensor[len(t), *y0.shape]``; default method ``dopri5``; fixed-grid options ``step_size`` /
``perturb``; adaptive options ``first_step, safety, ifactor, dfactor, max_num_steps``; unknown option keys warn;
solver failures raise ``AssertionError`` with torchdiffeq's messages.  Gradients are the discrete backprop through
the accepted steps (what autograd gives through torchdiffeq's ``odeint``), with constant step sizes.  One known
difference: torchdiffeq computes dopri5's FIRST step size from ``y0`` and ``f(t0, y0)`` inside the
graph, so its autograd gradient carries a term through that step size; the kernels treat every step size -- the first
included -- as a constant.  Measured against the oracle with the differentiable first step: 2e-6 .. 4e-5 norm-wise on
``dL/dy0`` / the encoder gradients at rtol 1e-7 (``tests/test_training_fixture.py``), bounded at 5e-3 by
``test_first_step_gradient_term_is_small`` for loose tolerances; pass ``options={'first_step': h}`` to remove the term
on both sides.

There is no generic path: ``func`` must be one of the hybrid-ODE vector fields (``TypeError`` otherwise) and ``y0``
must live on a CUDA device (``RuntimeError`` otherwise).  Extensions (all optional, in ``options``):
``controller='batch'|'trajectory'``, ``n_groups`` (several independent odeint calls in one launch),
``tape_capacity``, ``expert_grads``.
"""
from __future__ import annotations

import warnings
import weakref
from typing import Optional

import torch

from . import _lib as L
from . import ops

__all__ = ["odeint", "odeint_adjoint", "odeint_ensemble", "odeint_sse", "EnsembleParams", "SolveInfo", "last_solve_info",
           "last_adjoint_solve_info", "fixed_grid_points"]

EXPERT_NAMES = (
    "HillCure", "HillPatho", "ec50_patho", "emax_patho", "k_dexa", "k_discure_immunereact", "k_discure_immunity",
    "k_disprog", "k_immune_disease", "k_immune_feedback", "k_immune_off", "k_immunity", "kel",
)


class SolveInfo:
    """Counters of the most recent dopri5 solve (per controller): accepted, rejected, nfe, status."""

    def __init__(self, stats: Optional[torch.Tensor], n_steps_fixed: int = 0):
        self.stats = stats
        self.n_steps_fixed = n_steps_fixed

    @property
    def accepted(self):
        return None if self.stats is None else self.stats[:, 0]

    @property
    def rejected(self):
        return None if self.stats is None else self.stats[:, 1]

    @property
    def nfe(self):
        return None if self.stats is None else self.stats[:, 2]

    @property
    def attempts_total(self) -> int:
        if self.stats is None:
            return self.n_steps_fixed
        return int((self.stats[:, 0] + self.stats[:, 1]).sum())


_last_info: Optional[SolveInfo] = None


def last_solve_info() -> Optional[SolveInfo]:
    return _last_info


# ------------------------------------------------------------------------------------------------------------------
# recognising the vector field (ours, or the reference's own classes by structure) and packing its parameters
# ------------------------------------------------------------------------------------------------------------------
def _is_roche(func) -> bool:
    return all(hasattr(func, n) for n in EXPERT_NAMES) and hasattr(func, "ml_net") and hasattr(func, "dose_at_time")


def _is_neural(func) -> bool:
    return (
        hasattr(func, "ml_net") and hasattr(func, "kel") and not hasattr(func, "HillCure")
        and isinstance(func.ml_net, torch.nn.Sequential) and len(func.ml_net) == 4
    )


def field_kind(func):
    if isinstance(func, torch.nn.Module):
        if _is_roche(func):
            return L.FIELD_ROCHE
        if _is_neural(func):
            return L.FIELD_NEURAL
    raise TypeError(
        "hybrid_ode odeint only integrates the hybrid-ODE vector fields (RocheODE / NeuralODE); got {}. "
        "There is no generic fallback.".format(type(func).__name__)
    )


def packed_order(func, kind):
    """The vector field's parameters in the order of the packed layout of include/hode.h."""
    if kind == L.FIELD_ROCHE:
        parts = [getattr(func, n) for n in EXPERT_NAMES]
        if int(func.latent_dim) > 4:
            lin = func.ml_net[0]
            parts += [lin.weight, lin.bias]
        if getattr(func, "ablate", False):  # the ablation field's two scalars go last (include/hode.h)
            parts += [func.theta_1, func.theta_2]
    else:
        l1, l2 = func.ml_net[0], func.ml_net[2]
        parts = [func.kel, l1.weight, l1.bias, l2.weight, l2.bias]
    return parts


def pack_params(func, kind) -> torch.Tensor:
    """Differentiable packed parameter vector, layout of include/hode.h."""
    return torch.cat([p.reshape(-1) for p in packed_order(func, kind)]).to(torch.float32)


class EnsembleParams:
    """The parameters of ``M`` ensemble members / restarts as ONE ``[M, P]`` leaf tensor in the packed layout of
    include/hode.h (BASELINE config 4: ``run_simulation_ensemble.py:98-147`` trains the members one after the other).

    Every member's ``nn.Parameter`` is re-homed as a view of its row of ``flat`` (values preserved, ``state_dict`` keys
    unchanged), so ``odeint_ensemble(ens, ...)`` hands ``flat`` to the kernels as it is and autograd sees ONE leaf:
    the gradients of all members arrive in ``flat.grad`` straight from the backward kernel's ``grad_params[M, P]``.
    Without it a step packs ``M`` x 15 tensors and routes ``M`` x 15 gradient views through autograd -- at 64 members that
    host work is 6 x the kernels' time.  Train with an optimizer over ``[ens.flat]`` (element-wise optimizers such as
    Adam / SGD are unchanged by the flattening); ``ens.member_grad(m)`` maps a row of ``flat.grad`` back to names."""

    def __init__(self, funcs):
        funcs = list(funcs)
        if len(funcs) == 0:
            raise ValueError("EnsembleParams needs at least one member")
        kind = field_kind(funcs[0])
        if real_kind_of(funcs[0]) is not None:
            raise NotImplementedError("odeint_ensemble is built for the simulation fields only")
        self.funcs, self.kind = funcs, kind
        orders = [packed_order(f, kind) for f in funcs]
        sizes = [p.numel() for p in orders[0]]
        for f, o in zip(funcs, orders):
            if field_kind(f) != kind or [p.numel() for p in o] != sizes:
                raise ValueError("ensemble members must be vector fields of the same class and latent_dim")
        dev = orders[0][0].device
        flat = torch.empty(len(funcs), sum(sizes), device=dev, dtype=torch.float32)
        with torch.no_grad():
            for m, o in enumerate(orders):
                off = 0
                for p in o:
                    n = p.numel()
                    flat[m, off:off + n] = p.detach().reshape(-1).to(torch.float32)
                    p.data = flat[m, off:off + n].view(p.shape)  # the member keeps reading / loading its own tensors
                    off += n
        self.flat = torch.nn.Parameter(flat)
        self._sizes = sizes
        self._dose = None
        self._hill = None

    def __len__(self):
        return len(self.funcs)

    def parameters(self):
        return [self.flat]

    def set_action(self, action):
        """The same dose schedule for every member (the ensemble runs on one cohort): one ``set_action`` call."""
        f0 = self.funcs[0]
        f0.set_action(action)
        for f in self.funcs[1:]:
            f.dosage, f.times, f._dose_t_f32 = f0.dosage, f0.times, getattr(f0, "_dose_t_f32", None)
        self._dose = None

    def hill_exponents_are_two(self):
        """:func:`hill_exponents_are_two` for all members at once, cached per version of ``flat`` (one host read per
        optimizer step that touches the tensor; the kernels verify the flag on the device either way)."""
        if self.kind != L.FIELD_ROCHE:
            return False
        ver = self.flat._version
        if self._hill is None or self._hill[0] != ver:
            with torch.no_grad():
                self._hill = (ver, bool((self.flat[:, :2] == 2.0).all().item()))
        return self._hill[1]

    def member_grad(self, m):
        """``{position in the packed layout: gradient view}`` of member ``m`` after ``backward`` (views of ``flat.grad``)."""
        if self.flat.grad is None:
            return None
        out, off = [], 0
        for p, n in zip(packed_order(self.funcs[m], self.kind), self._sizes):
            out.append(self.flat.grad[m, off:off + n].view(p.shape))
            off += n
        return out


def real_kind_of(func):
    from .real import real_field_kind

    return real_field_kind(func)


_hill_cache = {}


def hill_exponents_are_two(func) -> bool:
    """True when ``HillCure == HillPatho == 2`` exactly (the ``RochConfig`` defaults; the simulation experiments never
    train them, ``run_simulation.py:125-129``): the kernels then compute ``x ** Hill`` as a multiply
    (``HODE_FLAG_HILL2``).  The two scalars live on the device, so the answer is cached per parameter version: one
    host read per change, none in a training loop that leaves the expert parameters alone."""
    hc, hp = func.HillCure, func.HillPatho
    key = (id(func), hc.data_ptr(), hp.data_ptr())
    ver = (hc._version, hp._version)
    hit = _hill_cache.get(key)
    if hit is not None and hit[0] == ver and hit[1]() is func:
        return hit[2]
    with torch.no_grad():
        ok = bool(((hc == 2.0) & (hp == 2.0)).item())
    if len(_hill_cache) > 256:
        _hill_cache.clear()
    _hill_cache[key] = (ver, weakref.ref(func), ok)
    return ok


def fixed_grid_points(t: torch.Tensor, step_size) -> torch.Tensor:
    """torchdiffeq's fixed grid, computed in ``t.dtype`` exactly as ``_grid_constructor_from_step_size`` does."""
    if step_size is None:
        return t
    start, end = t[0], t[-1]
    niters = torch.ceil((end - start) / step_size + 1).item()
    grid = torch.arange(0, niters, dtype=t.dtype, device=t.device) * step_size + start
    grid[-1] = t[-1]
    return grid


_time_cache = {}


def _times_for(t: torch.Tensor, step_size, device, fixed: bool):
    """Device copies of the output times / solver grid.  Cached per tensor OBJECT (weak reference + version counter),
    so a decoder's persistent ``self.t`` costs one host read in its lifetime; a fresh tensor is always re-read."""
    key = (id(t), step_size, str(device), fixed)
    hit = _time_cache.get(key)
    if hit is not None:
        ref, version, out = hit
        if ref() is t and version == t._version:
            return out
    t_host = t.detach().cpu()
    assert t_host.dim() == 1 and torch.is_floating_point(t_host), "t must be a one dimensional floating point Tensor"
    if t_host.numel() > 1:
        inc = bool((t_host[1:] > t_host[:-1]).all())
        dec = bool((t_host[1:] < t_host[:-1]).all())
        assert inc or dec, "t must be strictly increasing or decreasing"
        if dec:
            raise NotImplementedError("decreasing t (reverse-time solve) is not used by the reference and not built")
    if fixed:
        if t_host.dtype != torch.float32:
            raise NotImplementedError("fixed-grid solvers do their time arithmetic in t.dtype; only float32 t is built")
        grid = fixed_grid_points(t_host, step_size)
        assert grid[0] == t_host[0] and grid[-1] == t_host[-1]
        out = (t_host.to(device).contiguous(), grid.to(device).contiguous())
    else:
        out = (t_host.to(torch.float64).to(device).contiguous(), None)
    if len(_time_cache) > 64:
        _time_cache.clear()
    _time_cache[key] = (weakref.ref(t), t._version, out)
    return out


def adjoint_grid_points(t: torch.Tensor, step_size):
    """Solver grids of the continuous adjoint: torchdiffeq integrates interval ``i`` from ``t[i]`` back to ``t[i-1]`` by
    negating time (``odeint.py`` ``_check_inputs``), so the grid is the fixed grid of ``[-t[i], -t[i-1]]``.  Returns the
    grids back to back (last interval first) and the number of points of each."""
    parts, counts = [], []
    for i in range(t.numel() - 1, 0, -1):
        g = fixed_grid_points(-(t[i - 1:i + 1].flip(0)), step_size)
        assert g[0] == -t[i] and g[-1] == -t[i - 1]
        parts.append(g)
        counts.append(g.numel())
    grid = torch.cat(parts) if parts else torch.zeros(0, dtype=t.dtype)
    return grid, torch.tensor(counts, dtype=torch.int32)


_adj_time_cache = {}


def _adjoint_times_for(t: torch.Tensor, step_size, device):
    key = (id(t), step_size, str(device))
    hit = _adj_time_cache.get(key)
    if hit is not None:
        ref, version, out = hit
        if ref() is t and version == t._version:
            return out
    grid, counts = adjoint_grid_points(t.detach().cpu(), step_size)
    out = (grid.to(device).contiguous(), counts.to(device).contiguous())
    if len(_adj_time_cache) > 64:
        _adj_time_cache.clear()
    _adj_time_cache[key] = (weakref.ref(t), t._version, out)
    return out


# ------------------------------------------------------------------------------------------------------------------
class _FixedSolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, packed, pb, grid, t_eval, need_grad):
        lib = L.get_lib()
        pb.params = packed.detach().reshape(pb.params_shape).contiguous()
        h, tape = ops.fixed_fwd(lib, pb, y0.detach(), grid, t_eval, need_grad)
        ctx.pb, ctx.grid, ctx.t_eval, ctx.tape = pb, grid, t_eval, tape
        return h

    @staticmethod
    def backward(ctx, grad_h):
        if ctx.tape is None:
            raise RuntimeError("backward through a solve that was run without a tape")
        gy0, gp = ops.fixed_bwd(L.get_lib(), ctx.pb, ctx.grid, ctx.t_eval, grad_h, ctx.tape)
        return gy0, gp.reshape(-1), None, None, None, None


class _FixedSolveSSE(torch.autograd.Function):
    """Forward solve with the read-out and the masked SSE consumed at every output time (``hode_fixed_fwd_sse``): returns
    the scalar loss; ``d loss / d h`` is produced in the same launch and fed to the reverse sweep in ``backward``."""

    @staticmethod
    def forward(ctx, y0, packed, W, b, pb, grid, t_eval, x, mask, n_norm, need_grad):
        lib = L.get_lib()
        pb.params = packed.detach().reshape(pb.params_shape).contiguous()
        loss, gh, gw, gb, _, tape = ops.fixed_fwd_sse(lib, pb, y0.detach(), grid, t_eval, W.detach(), b.detach(), x, mask, n_norm,
                                                      want_tape=need_grad, want_param_grads=need_grad)
        ctx.pb, ctx.grid, ctx.t_eval, ctx.tape = pb, grid, t_eval, tape
        if need_grad:
            ctx.save_for_backward(gh, gw, gb)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        if ctx.tape is None:
            raise RuntimeError("backward through a solve that was run without a tape")
        gh, gw, gb = ctx.saved_tensors
        gy0, gp = ops.fixed_bwd(L.get_lib(), ctx.pb, ctx.grid, ctx.t_eval, gh, ctx.tape)
        # the loss is a scalar: its upstream gradient scales the (small) results instead of the [n_t, n_traj, D] grad_h
        return g * gy0, g * gp.reshape(-1), g * gw, g * gb, None, None, None, None, None, None, None


class _FixedAdjointSolve(torch.autograd.Function):
    """torchdiffeq's ``OdeintAdjointMethod`` for the fixed-grid methods: forward without a tape, backward = ONE launch
    integrating the augmented system backwards over every output interval (``hode_fixed_adjoint``)."""

    @staticmethod
    def forward(ctx, y0, packed, pb, grid, t_eval, adj_pb, adj_grid, adj_count):
        lib = L.get_lib()
        pb.params = packed.detach().reshape(pb.params_shape).contiguous()
        adj_pb.params = pb.params
        h, _ = ops.fixed_fwd(lib, pb, y0.detach(), grid, t_eval, False)
        ctx.adj_pb, ctx.adj_grid, ctx.adj_count = adj_pb, adj_grid, adj_count
        ctx.save_for_backward(h)
        return h

    @staticmethod
    def backward(ctx, grad_h):
        (h,) = ctx.saved_tensors
        gy0, gp = ops.fixed_adjoint(L.get_lib(), ctx.adj_pb, ctx.adj_grid, ctx.adj_count, h, grad_h)
        return gy0, gp.reshape(-1), None, None, None, None, None, None


class _Dopri5Solve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, packed, pb, t_eval, tape_capacity, need_grad, holder):
        lib = L.get_lib()
        pb.params = packed.detach().reshape(pb.params_shape).contiguous()
        cap = int(tape_capacity) if need_grad else 0
        while True:
            h, stats, tape = ops.dopri5_fwd(lib, pb, y0.detach(), t_eval, cap)
            st = stats.cpu()
            status = st[:, 3]
            if need_grad and bool((status == L.SOLVE_TAPE_FULL).any()) and cap < (1 << 20):
                cap *= 4  # the tape was too short for this solve: rerun with a larger one
                continue
            break
        holder.append(st)
        _raise_on_failure(st)
        ctx.pb, ctx.t_eval, ctx.tape, ctx.stats = pb, t_eval, tape, stats
        return h

    @staticmethod
    def backward(ctx, grad_h):
        if ctx.tape is None:
            raise RuntimeError("backward through a solve that was run without a tape")
        gy0, gp = ops.dopri5_bwd(L.get_lib(), ctx.pb, ctx.t_eval, grad_h, ctx.tape, ctx.stats)
        return gy0, gp.reshape(-1), None, None, None, None, None


class _Dopri5AdjointSolve(torch.autograd.Function):
    """torchdiffeq's ``OdeintAdjointMethod`` with the adaptive solver: forward = the dopri5 solve without a tape, backward = ONE
    launch integrating the augmented system backwards over every output interval with the dopri5 controller
    (``hode_dopri5_adjoint``; error control by torchdiffeq's mixed norm or by 'seminorm')."""

    @staticmethod
    def forward(ctx, y0, packed, pb, adj_pb, t_eval, holder):
        lib = L.get_lib()
        pb.params = packed.detach().reshape(pb.params_shape).contiguous()
        adj_pb.params = pb.params
        h, stats, _ = ops.dopri5_fwd(lib, pb, y0.detach(), t_eval, 0)
        st = stats.cpu()
        holder.append(st)
        _raise_on_failure(st)
        ctx.adj_pb, ctx.t_eval = adj_pb, t_eval
        ctx.save_for_backward(h)
        return h

    @staticmethod
    def backward(ctx, grad_h):
        (h,) = ctx.saved_tensors
        global _last_adjoint_info
        gy0, gp, stats = ops.dopri5_adjoint(L.get_lib(), ctx.adj_pb, ctx.t_eval, h, grad_h)
        st = stats.cpu()
        _last_adjoint_info = SolveInfo(st)
        try:
            _raise_on_failure(st)
        except AssertionError as e:
            cfg = ctx.adj_pb.cfg
            if not (cfg.flags & L.FLAG_ADJ_SEMINORM) and cfg.expert_grads:
                # torchdiffeq fails the same way ('underflow in dt nan'): under its default mixed norm ONE NaN in a parameter's
                # error estimate -- d f / d Hill = x**p log x at a state an attempt pushed below zero -- makes the ratio NaN
                raise AssertionError(
                    str(e) + " -- adjoint solve under torchdiffeq's default mixed norm with the 13 expert scalars among the "
                    "adjoint parameters: a NaN in d f / d HillCure / d HillPatho poisons the error norm (torchdiffeq stops the "
                    "same way).  Pass options={'expert_grads': False} (the reference never trains the expert scalars, "
                    "run_simulation.py:125-129) or adjoint_options={'norm': 'seminorm'}") from None
            raise
        return gy0, gp.reshape(-1), None, None, None, None


_last_adjoint_info = None


def last_adjoint_solve_info():
    """Counters of the most recent ADAPTIVE ADJOINT solve (``odeint_adjoint(method='dopri5')`` backward pass): accepted /
    rejected attempts per controller summed over the output intervals, status."""
    return _last_adjoint_info


def _raise_on_failure(st: torch.Tensor):
    status = st[:, 3]
    if bool((status == L.SOLVE_OK).all()):
        return
    bad = int(torch.nonzero(status != L.SOLVE_OK)[0])
    code = int(status[bad])
    # torchdiffeq raises these with `assert` (rk_common.py); callers such as training_utils.py:45 rely on the type
    if code == L.SOLVE_DT_UNDERFLOW:
        raise AssertionError("underflow in dt (controller {})".format(bad))
    if code == L.SOLVE_NONFINITE:
        raise AssertionError("non-finite values in state `y` (controller {})".format(bad))
    if code == L.SOLVE_MAX_STEPS:
        raise AssertionError("max_num_steps exceeded (controller {})".format(bad))
    raise RuntimeError("dopri5 tape capacity exceeded (controller {})".format(bad))


_FIXED_KEYS = {"step_size", "perturb", "grid_constructor", "interp"}
_ADAPTIVE_KEYS = {"first_step", "step_t", "jump_t", "safety", "ifactor", "dfactor", "max_num_steps", "dtype", "norm"}
_OUR_KEYS = {"controller", "n_groups", "tape_capacity", "expert_grads", "attempt_cap", "param_sets", "param_set_of_group",
             "hill2_kernels"}
_NAMES = {"euler": "Euler", "midpoint": "Midpoint", "rk4": "RK4", "dopri5": "Dopri5Solver"}


def _dose_tensors(func, n, device):
    if func.dosage is None or func.times is None:
        raise RuntimeError("set_action must be called before integrating (model.py:1113)")
    dose_amt = func.dosage.detach().to(device=device, dtype=torch.float32).contiguous()
    dose_t = getattr(func, "_dose_t_f32", None)
    if dose_t is None or dose_t.shape[0] != n:
        dose_t = func.times.detach().to(device=device, dtype=torch.float32).contiguous()
    if dose_amt.shape[0] != n or dose_t.shape[0] != n:
        raise RuntimeError("dose schedule has {} patients but y0 has {}".format(dose_amt.shape[0], n))
    return dose_amt, dose_t


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    """torchdiffeq's ``odeint`` for one vector field (module docstring)."""
    return _odeint_impl([func], y0, t, rtol, atol, method, options, event_fn)


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None, adjoint_rtol=None,
                   adjoint_atol=None, adjoint_method=None, adjoint_options=None, adjoint_params=None):
    """torchdiffeq's ``odeint_adjoint`` (the import the reference keeps commented out at ``model.py:9``): same forward
    result as :func:`odeint`, but the backward pass is the CONTINUOUS adjoint -- no tape, O(1) memory in the number of
    solver steps.  Fixed-grid methods (``euler`` / ``midpoint`` / ``rk4``) and ``dopri5`` (the reference's default method;
    error control of the adaptive adjoint solve by torchdiffeq's default mixed norm -- RocheODE up to latent_dim 8, batch-coupled
    controller -- or by ``adjoint_options={'norm': 'seminorm'}`` -- every field and controller; tolerances ``adjoint_rtol`` /
    ``adjoint_atol`` defaulting to the forward ones); the adjoint solve uses the forward method and options unless ``adjoint_method`` /
    ``adjoint_options`` say otherwise (torchdiffeq's defaults).
    ``adjoint_params`` must be the vector field's own parameters (the default): gradients are produced for the packed
    parameter vector as a whole.  Gradients differ from :func:`odeint`'s discrete backprop by the method's
    discretisation error, exactly as they do in torchdiffeq."""
    if adjoint_params is not None:
        own = {id(p) for p in func.parameters()}
        if {id(p) for p in adjoint_params} - own:
            raise NotImplementedError("adjoint_params must be parameters of the vector field")
    method = "dopri5" if method is None else method
    adjoint_method = method if adjoint_method is None else adjoint_method
    for m in (method, adjoint_method):
        if m not in L.METHODS:
            raise ValueError('Invalid method "{}". Must be one of {}'.format(m, "{" + ", ".join(L.METHODS) + "}"))
    if (method == "dopri5") != (adjoint_method == "dopri5"):
        raise NotImplementedError("odeint_adjoint: the forward and the adjoint solve must both be fixed-grid or both dopri5")
    if adjoint_options is None:  # torchdiffeq: the forward options (minus a user norm) drive the adjoint solve too
        adjoint_options = {k: v for k, v in options.items() if k != "norm"} if options is not None else {}
    else:
        adjoint_options = dict(adjoint_options)
        known = (_FIXED_KEYS if adjoint_method != "dopri5" else _ADAPTIVE_KEYS) | _OUR_KEYS
        unused = {k: v for k, v in adjoint_options.items() if k not in known}
        if unused:
            warnings.warn("{}: Unexpected arguments {}".format(_NAMES[adjoint_method], unused))
    if adjoint_method == "dopri5":
        norm = adjoint_options.get("norm")
        if norm is not None and norm != "seminorm":
            raise NotImplementedError("adjoint_options['norm'] must be absent (torchdiffeq's mixed norm) or 'seminorm'")
        if norm is None:
            # torchdiffeq's default: the MIXED norm -- every parameter tensor's adjoint takes part in the error control
            # (hode_dopri5_adjoint without HODE_FLAG_ADJ_SEMINORM: batch-coupled controller, RocheODE up to latent_dim 8)
            from .real import real_field_kind
            opts = options or {}
            ctrl = adjoint_options.get("controller", opts.get("controller", "batch"))
            if (real_field_kind(func) is not None or field_kind(func) != L.FIELD_ROCHE or int(func.latent_dim) > 8
                    or ctrl != "batch"):
                raise NotImplementedError(
                    "odeint_adjoint(method='dopri5') with torchdiffeq's default mixed norm (every parameter adjoint can veto a "
                    "step) is built for the RocheODE field up to latent_dim 8 with the batch-coupled controller; pass "
                    "adjoint_options={'norm': 'seminorm'} (error control by the state and the state adjoint) otherwise")
    adjoint = {"method": adjoint_method, "options": adjoint_options,
               "rtol": rtol if adjoint_rtol is None else adjoint_rtol, "atol": atol if adjoint_atol is None else adjoint_atol}
    return _odeint_impl([func], y0, t, rtol, atol, method, options, event_fn, adjoint=adjoint)


def odeint_ensemble(funcs, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None):
    """``M`` independent ``odeint`` calls -- one per ensemble member / restart, each with its OWN parameters -- in ONE
    kernel launch (BASELINE config 4; the reference runs restarts and methods one after the other:
    ``run_simulation.py:95``, ``Fig3.sh:12-50``).

    ``funcs``: ``M`` vector fields of the same class and ``latent_dim``, each after its own ``set_action`` on ``B``
    patients -- or an :class:`EnsembleParams` over them (all members' parameters as one ``[M, P]`` leaf: no per-member
    packing, one gradient tensor; the path to use beyond a handful of members).  ``y0``: ``[M * B, D]``, member-major.  Returns ``[len(t), M * B, D]``; gradients flow to every member's
    parameters.  Each member is its own controller group (``options={'n_groups': g}`` splits a member into ``g`` groups).
    """
    if isinstance(funcs, EnsembleParams):
        return _odeint_impl(funcs, y0, t, rtol, atol, method, options, None)
    funcs = list(funcs)
    if len(funcs) == 0:
        raise ValueError("odeint_ensemble needs at least one member")
    return _odeint_impl(funcs, y0, t, rtol, atol, method, options, None)


def odeint_sse(func, y0, t, weight, bias, x, mask, *, n_norm=None, rtol=1e-7, atol=1e-9, method=None, options=None):
    """The likelihood term of ``VariationalInference.loss`` in one call: ``h = odeint(func, y0, t, ...)`` (``model.py:1116``),
    ``x_hat = h @ weight.T + bias`` (``model.py:1120``) and ``sum((x - x_hat)**2 * mask) / n_norm`` (``model.py:1179``,
    ``n_norm = x.shape[1]`` by default), returned as a scalar that back-propagates to ``y0``, the vector-field parameters,
    ``weight`` and ``bias``.

    For the fixed-grid methods on the RocheODE field this is ONE forward launch (the read-out and the loss are evaluated
    while ``h(t_j)`` is in registers; ``h`` and ``x_hat`` never exist in memory) plus the reverse sweep in ``backward``.  Where
    the fused kernel does not apply (dopri5, NeuralODE, several parameter sets, ``obs % 4 != 0``, strided data) the same
    result comes from ``odeint`` + the fused read-out / SSE kernel (``loss.decode_sse_loss``) -- also CUDA, two launches more."""
    global _last_info
    from .loss import decode_sse_loss
    from .real import real_field_kind

    n_norm = float(x.shape[1] if n_norm is None else n_norm)
    fusable = real_field_kind(func) is None and (method is not None and method != "dopri5")
    if fusable:
        su = _setup([func], y0, t, rtol, atol, method, options, None)
        opts = su["options"]
        xm = x if x.dtype == torch.float32 else x.float()
        mm = mask if mask.dtype == torch.float32 else mask.float()
        if xm.stride() != mm.stride() or not xm.is_contiguous():
            xm, mm = xm.contiguous(), mm.contiguous()
        pb = su["pb"]
        pb.params = su["packed"].detach().reshape(pb.params_shape)
        if (opts.get("grid_constructor") is None and opts.get("interp", "linear") == "linear"
                and ops.fixed_fwd_sse_supported(L.get_lib(), pb, weight.shape[0], xm, mm)):
            t_dev, grid = _times_for(t, opts.get("step_size"), y0.device, True)
            loss = _FixedSolveSSE.apply(y0, su["packed"], weight, bias, pb, grid, t_dev, xm, mm, n_norm, su["need_grad"] or
                                        (torch.is_grad_enabled() and (weight.requires_grad or bias.requires_grad)))
            _last_info = SolveInfo(None, (grid.numel() - 1) * su["B"])
            return loss
    h = odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
    return decode_sse_loss(h, weight, bias, x, mask, n_norm)


def _odeint_real(func, kind, y0, t, method, options):
    """Real-data fields (model.py:570-769): fixed-grid solvers only, one parameter set, dose tables built per launch."""
    global _last_info
    from . import real as _real

    if not isinstance(y0, torch.Tensor) or not y0.is_cuda:
        raise RuntimeError("hybrid_ode odeint runs on CUDA (sm_100a) only. There is no CPU fallback.")
    if y0.dtype != torch.float32 or y0.dim() != 2:
        raise NotImplementedError("real-data fields integrate float32 states of shape [batch, latent_dim]")
    method = "dopri5" if method is None else method
    if method not in L.METHODS:
        raise ValueError('Invalid method "{}". Must be one of {}'.format(method, "{" + ", ".join(L.METHODS) + "}"))
    if method == "dopri5":
        raise NotImplementedError("the real-data fields are integrated with euler / midpoint / rk4 only "
                                  "(experiments/real.sh:9-17); dopri5 on them has no fused kernel")
    options = {} if options is None else dict(options)
    unused = {k: v for k, v in options.items() if k not in (_FIXED_KEYS | _OUR_KEYS)}
    if unused:
        warnings.warn("{}: Unexpected arguments {}".format(_NAMES[method], unused))
    if y0.shape[1] != int(func.latent_dim):
        raise ValueError("y0 has {} columns but the vector field has latent_dim {}".format(y0.shape[1], func.latent_dim))
    need_grad = torch.is_grad_enabled() and (y0.requires_grad or any(p.requires_grad for p in func.parameters()))
    t_dev, grid = _times_for(t, options.get("step_size"), y0.device, True)
    h = _real.solve_real(func, kind, y0, method, bool(options.get("perturb", False)), grid, t_dev, need_grad)
    _last_info = SolveInfo(None, (grid.numel() - 1) * y0.shape[0])
    return h


def _setup(funcs, y0, t, rtol, atol, method, options, event_fn):
    """Argument checking shared by every entry point for the simulation fields; returns the launch description."""
    ens = funcs if isinstance(funcs, EnsembleParams) else None
    if ens is not None:
        funcs = ens.funcs
    func = funcs[0]
    M = len(funcs)
    if event_fn is not None:
        raise NotImplementedError("event_fn is not used by the reference and has no fused kernel")
    kind = field_kind(func)
    if ens is None:
        for f in funcs[1:]:
            if field_kind(f) != kind or int(f.latent_dim) != int(func.latent_dim):
                raise ValueError("ensemble members must be vector fields of the same class and latent_dim")
    if not isinstance(y0, torch.Tensor):
        raise NotImplementedError("tuple states are not used by the reference and have no fused kernel")
    if not y0.is_cuda:
        raise RuntimeError(
            "hybrid_ode odeint runs on CUDA (sm_100a) only; y0 is on {}. There is no CPU fallback.".format(y0.device)
        )
    if y0.dtype != torch.float32:
        raise NotImplementedError("only float32 states (the reference's DTYPE, global_config.py:3) have kernels")
    assert y0.dim() == 2, "y0 must be [batch, latent_dim]"
    for name, tol in (("rtol", rtol), ("atol", atol)):
        if isinstance(tol, torch.Tensor):
            assert not tol.requires_grad, name + " cannot require gradient"
    method = "dopri5" if method is None else method
    if method not in L.METHODS:
        raise ValueError('Invalid method "{}". Must be one of {}'.format(method, "{" + ", ".join(L.METHODS) + "}"))
    options = {} if options is None else dict(options)
    fixed = method != "dopri5"
    known = (_FIXED_KEYS if fixed else _ADAPTIVE_KEYS) | _OUR_KEYS
    unused = {k: v for k, v in options.items() if k not in known}
    if unused:
        warnings.warn("{}: Unexpected arguments {}".format(_NAMES[method], unused))

    B, D = y0.shape
    if D != int(func.latent_dim):
        raise ValueError("y0 has {} columns but the vector field has latent_dim {}".format(D, func.latent_dim))
    if B % M != 0:
        raise ValueError("y0 has {} rows, not a multiple of the {} ensemble members".format(B, M))
    groups_per_member = int(options.get("n_groups", 1))
    n_groups = groups_per_member * M
    if groups_per_member < 1 or (B // M) % groups_per_member != 0:
        raise ValueError("n_groups must divide the batch")
    batch = B // n_groups
    dose_key = None if ens is None else tuple((id(f.dosage), id(f.times)) for f in funcs)
    if M == 1:
        dose_amt, dose_t = _dose_tensors(func, B, y0.device)
    elif ens is not None and ens._dose is not None and ens._dose[0] == dose_key and ens._dose[1].shape[0] == B:
        _, dose_amt, dose_t = ens._dose  # the members' schedules have not been re-assigned since the last call
    else:
        parts = [_dose_tensors(f, B // M, y0.device) for f in funcs]
        if len({p[1].shape[1:] for p in parts}) != 1:
            raise RuntimeError("ensemble members must have the same number of doses per patient")
        dose_amt = torch.cat([p[0] for p in parts]).contiguous()
        dose_t = torch.cat([p[1] for p in parts]).contiguous()
        if ens is not None:
            ens._dose = (dose_key, dose_amt, dose_t)
    n_dose = dose_t.shape[1] if dose_t.dim() == 2 else 0

    ctrl_name = options.get("controller", "batch")
    if ctrl_name not in ("batch", "trajectory"):
        raise ValueError("controller must be 'batch' or 'trajectory'")
    need_grad = torch.is_grad_enabled() and (
        y0.requires_grad or (ens.flat.requires_grad if ens is not None else
                             any(p.requires_grad for f in funcs for p in f.parameters()))
    )
    ablate = kind == L.FIELD_ROCHE and bool(getattr(func, "ablate", False))
    if kind == L.FIELD_ROCHE and any(bool(getattr(f, "ablate", False)) != ablate for f in funcs):
        raise ValueError("ensemble members must all be ablation fields or none")
    cfg = ops.make_cfg(
        kind, D, L.METHODS[method],
        controller=L.CTRL_TRAJ if ctrl_name == "trajectory" else L.CTRL_BATCH,
        perturb=bool(options.get("perturb", False)), n_dose=n_dose,
        expert_grads=bool(options.get("expert_grads", True)),
        rtol=float(rtol), atol=float(atol), safety=float(options.get("safety", 0.9)),
        ifactor=float(options.get("ifactor", 10.0)), dfactor=float(options.get("dfactor", 0.2)),
        first_step=options.get("first_step", None), max_num_steps=int(options.get("max_num_steps", 2 ** 31 - 1)),
        attempt_cap=int(options.get("attempt_cap", ops.ATTEMPT_CAP_DEFAULT)),
        hill2=(kind == L.FIELD_ROCHE and bool(options.get("hill2_kernels", True))
               and (ens.hill_exponents_are_two() if ens is not None else all(hill_exponents_are_two(f) for f in funcs))),
        ablate=ablate,
    )
    if M == 1:
        packed = pack_params(func, kind)
        pset = None
    else:
        packed = ens.flat.reshape(-1) if ens is not None else torch.stack([pack_params(f, kind) for f in funcs]).reshape(-1)
        pset = torch.arange(M, dtype=torch.int32, device=y0.device).repeat_interleave(groups_per_member).contiguous()
    pb = ops.Problem(cfg, n_groups, batch, dose_amt, dose_t, None, pset)
    pb.params_shape = (M, packed.numel() // M)
    return dict(pb=pb, cfg=cfg, packed=packed, pset=pset, method=method, options=options, fixed=fixed, need_grad=need_grad,
                ctrl_name=ctrl_name, B=B, D=D, batch=batch, n_groups=n_groups, dose_amt=dose_amt, dose_t=dose_t)


def _odeint_impl(funcs, y0, t, rtol, atol, method, options, event_fn, adjoint=None):
    global _last_info
    func = funcs.funcs[0] if isinstance(funcs, EnsembleParams) else funcs[0]
    M = len(funcs)
    from .real import real_field_kind

    rk = real_field_kind(func)
    if rk is not None:
        if event_fn is not None:
            raise NotImplementedError("event_fn is not used by the reference and has no fused kernel")
        if M != 1:
            raise NotImplementedError("odeint_ensemble is built for the simulation fields only")
        if adjoint is not None:
            raise NotImplementedError("odeint_adjoint is built for the simulation fields (RocheODE / NeuralODE) only")
        return _odeint_real(func, rk, y0, t, method, options)
    su = _setup(funcs, y0, t, rtol, atol, method, options, event_fn)
    pb, cfg, packed, pset, method, options = su["pb"], su["cfg"], su["packed"], su["pset"], su["method"], su["options"]
    fixed, need_grad, ctrl_name, B, batch, n_groups = su["fixed"], su["need_grad"], su["ctrl_name"], su["B"], su["batch"], su["n_groups"]
    dose_amt, dose_t = su["dose_amt"], su["dose_t"]

    if fixed:
        if options.get("grid_constructor") is not None:
            raise NotImplementedError("grid_constructor is not used by the reference; pass step_size")
        if options.get("interp", "linear") != "linear":
            raise ValueError("Unknown interpolation method {}".format(options.get("interp")))
        t_dev, grid = _times_for(t, options.get("step_size"), y0.device, True)
        if adjoint is not None and need_grad:
            aopt = adjoint["options"]
            if aopt.get("grid_constructor") is not None or aopt.get("interp", "linear") != "linear":
                raise NotImplementedError("adjoint_options: only step_size / perturb are built")
            acfg = L.HodeCfg.from_buffer_copy(cfg)
            acfg.method, acfg.perturb = L.METHODS[adjoint["method"]], int(bool(aopt.get("perturb", False)))
            adj_pb = ops.Problem(acfg, n_groups, batch, dose_amt, dose_t, None, pset)
            adj_grid, adj_count = _adjoint_times_for(t, aopt.get("step_size"), y0.device)
            h = _FixedAdjointSolve.apply(y0, packed, pb, grid, t_dev, adj_pb, adj_grid, adj_count)
        else:
            h = _FixedSolve.apply(y0, packed, pb, grid, t_dev, need_grad)
        _last_info = SolveInfo(None, (grid.numel() - 1) * B)
        return h

    for key in ("step_t", "jump_t"):
        v = options.get(key)
        if v is not None and len(v) > 0:
            raise NotImplementedError("dopri5 option {} has no fused kernel (the reference only passes it to fixed-grid "
                                      "solvers, which ignore it)".format(key))
    if ctrl_name == "batch":
        mb = int(L.get_lib().hode_dopri5_max_batch(cfg))
        if batch > mb:
            raise NotImplementedError(
                "batch-coupled dopri5 integrates one group per CTA or cluster of CTAs (<= {} trajectories); got {}. Use "
                "options={{'n_groups': g}} or options={{'controller': 'trajectory'}}.".format(mb, batch)
            )
    t_dev, _ = _times_for(t, None, y0.device, False)
    holder = []
    if adjoint is not None and need_grad:
        if ctrl_name == "batch" and batch > 512:
            raise NotImplementedError("the adaptive adjoint integrates one group per CTA (<= 512 trajectories); got {}. Use "
                                      "options={{'n_groups': g}} or options={{'controller': 'trajectory'}}.".format(batch))
        aopt = adjoint["options"]
        acfg = L.HodeCfg.from_buffer_copy(cfg)
        acfg.rtol, acfg.atol = float(adjoint["rtol"]), float(adjoint["atol"])
        if aopt.get("norm") == "seminorm":
            acfg.flags |= L.FLAG_ADJ_SEMINORM
        for key, attr in (("safety", "safety"), ("ifactor", "ifactor"), ("dfactor", "dfactor")):
            if key in aopt:
                setattr(acfg, attr, float(aopt[key]))
        acfg.first_step = float(aopt["first_step"]) if aopt.get("first_step") is not None else -1.0
        if "max_num_steps" in aopt:
            acfg.max_num_steps = int(aopt["max_num_steps"])
        if "controller" in aopt:
            acfg.controller = L.CTRL_TRAJ if aopt["controller"] == "trajectory" else L.CTRL_BATCH
        adj_pb = ops.Problem(acfg, n_groups, batch, dose_amt, dose_t, None, pset)
        h = _Dopri5AdjointSolve.apply(y0, packed, pb, adj_pb, t_dev, holder)
        _last_info = SolveInfo(holder[0] if holder else None)
        return h
    h = _Dopri5Solve.apply(y0, packed, pb, t_dev, int(options.get("tape_capacity", 1024)), need_grad, holder)
    _last_info = SolveInfo(holder[0] if holder else None)
    return h
