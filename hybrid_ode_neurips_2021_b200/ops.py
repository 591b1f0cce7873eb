"""Thin tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers, check return codes.

Every function takes the :class:`HodeLib` to call, so the same wrappers drive the CUDA library in the product and the
test-only host emulation in ``tests/`` (CPU tensors).  No arithmetic of the hot path happens in this file.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib as L

ATTEMPT_CAP_DEFAULT = 1 << 22


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    if t.is_cuda:
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return None


def _on(t: torch.Tensor):
    """Device guard for one C-ABI call: the library launches on the CURRENT device (its constant-bank lease and every kernel
    launch use it), so a tensor living on another GPU -- the reference's default is ``cuda:1`` without ``set_device``
    (``global_config.py:7``, ``run_simulation.py:33``) -- must make its device current for the duration of the call."""
    if t.is_cuda:
        return torch.cuda.device(t.device)
    return contextlib.nullcontext()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float32, "float32 expected, got {}".format(t.dtype)
    return t.contiguous()


def make_cfg(field, latent_dim, method, *, controller=L.CTRL_BATCH, perturb=False, n_dose=1, expert_grads=True,
             rtol=1e-7, atol=1e-9, safety=0.9, ifactor=10.0, dfactor=0.2, first_step=None,
             max_num_steps=2 ** 31 - 1, attempt_cap=ATTEMPT_CAP_DEFAULT, hill2=False, ablate=False, adj_seminorm=False) -> L.HodeCfg:
    cfg = L.HodeCfg()
    cfg.flags = (L.FLAG_HILL2 if hill2 else 0) | (L.FLAG_ABLATE if ablate else 0) | (L.FLAG_ADJ_SEMINORM if adj_seminorm else 0)
    cfg.field, cfg.latent_dim, cfg.method, cfg.controller = int(field), int(latent_dim), int(method), int(controller)
    cfg.perturb, cfg.n_dose, cfg.expert_grads = int(bool(perturb)), int(n_dose), int(bool(expert_grads))
    cfg.rtol, cfg.atol, cfg.safety, cfg.ifactor, cfg.dfactor = float(rtol), float(atol), float(safety), float(ifactor), float(dfactor)
    cfg.first_step = -1.0 if first_step is None else float(first_step)
    cfg.max_num_steps, cfg.attempt_cap = int(max_num_steps), int(attempt_cap)
    return cfg


def dose_schedule(lib, action: torch.Tensor):
    """``action [T, B, 1]`` (any strides) -> ``dose_amt [B]`` f32, ``dose_idx [B, T]`` i32, ``dose_count [B]`` i32."""
    assert action.dim() == 3 and action.shape[2] == 1 and action.dtype == torch.float32
    T, B = action.shape[0], action.shape[1]
    dev = action.device
    amt = torch.empty(B, dtype=torch.float32, device=dev)
    idx = torch.empty(B, T, dtype=torch.int32, device=dev)
    cnt = torch.empty(B, dtype=torch.int32, device=dev)
    with _on(action):
        rc = lib.hode_dose_schedule(_ptr(action), action.stride(0), action.stride(1), T, B, _ptr(amt), _ptr(idx), _ptr(cnt),
                                    _stream(action))
    lib.check(rc, "hode_dose_schedule")
    return amt, idx, cnt


@dataclass
class Problem:
    """Device-resident inputs shared by forward and backward of one solve."""

    cfg: L.HodeCfg
    n_groups: int
    batch: int
    dose_amt: torch.Tensor  # [n_traj] f32
    dose_t: torch.Tensor  # [n_traj, stride] f32
    params: torch.Tensor  # [n_sets, P] f32
    pset: Optional[torch.Tensor]  # [n_groups] i32 or None

    @property
    def n_traj(self):
        return self.n_groups * self.batch


def fixed_fwd(lib, pb: Problem, y0, grid, t_eval, want_tape):
    D = pb.cfg.latent_dim
    y0 = _f32c(y0)
    n_t, n_grid = t_eval.numel(), grid.numel()
    h = torch.empty(n_t, pb.n_traj, D, dtype=torch.float32, device=y0.device)
    tape = torch.empty(max(n_grid - 1, 0), pb.n_traj, D, dtype=torch.float32, device=y0.device) if want_tape else None
    with _on(y0):
        rc = lib.hode_fixed_fwd(C.byref(pb.cfg), pb.n_groups, pb.batch, _ptr(y0), _ptr(pb.dose_amt), _ptr(pb.dose_t),
                                pb.dose_t.stride(0), _ptr(pb.params), _ptr(pb.pset), _ptr(grid), n_grid, _ptr(t_eval), n_t,
                                _ptr(h), _ptr(tape), _stream(y0))
    lib.check(rc, "hode_fixed_fwd")
    return h, tape


def fixed_fwd_sse_supported(lib, pb: Problem, obs, x=None, mask=None) -> bool:
    """Is there a fused solve + read-out + masked-SSE kernel for this problem (and are x / mask laid out for it)?"""
    if pb.pset is not None or pb.params.shape[0] != 1:
        return False
    if not lib.hode_fixed_fwd_sse_supported(C.byref(pb.cfg), int(obs), int(pb.params.shape[0])):
        return False
    for t in (x, mask):
        if t is not None and not (t.dtype == torch.float32 and t.is_contiguous() and t.data_ptr() % 16 == 0):
            return False
    return True


def fixed_fwd_sse(lib, pb: Problem, y0, grid, t_eval, W, b, x, mask, n_norm, want_tape, want_h=False, want_param_grads=True):
    """``hode_fixed_fwd_sse``: forward solve with read-out + masked SSE consumed at every output time.
    Returns ``loss [1], grad_h [n_t, n_traj, D], grad_W, grad_b, h or None, tape or None``."""
    D = pb.cfg.latent_dim
    y0, W, b = _f32c(y0), _f32c(W), _f32c(b)
    n_traj = pb.n_traj
    n_t, n_grid, obs = t_eval.numel(), grid.numel(), W.shape[0]
    assert x.shape == (n_t, n_traj, obs) and mask.shape == x.shape and x.is_contiguous() and mask.is_contiguous()
    dev = y0.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    gh = torch.empty(n_t, n_traj, D, dtype=torch.float32, device=dev)
    gw = torch.empty(obs, D, dtype=torch.float32, device=dev) if want_param_grads else None
    gb = torch.empty(obs, dtype=torch.float32, device=dev) if want_param_grads else None
    h = torch.empty(n_t, n_traj, D, dtype=torch.float32, device=dev) if want_h else None
    tape = torch.empty(max(n_grid - 1, 0), n_traj, D, dtype=torch.float32, device=dev) if want_tape else None
    with _on(y0):
        rc = lib.hode_fixed_fwd_sse(C.byref(pb.cfg), n_traj, _ptr(y0), _ptr(pb.dose_amt), _ptr(pb.dose_t), pb.dose_t.stride(0),
                                    _ptr(pb.params), _ptr(grid), n_grid, _ptr(t_eval), n_t, _ptr(W), _ptr(b), obs, _ptr(x),
                                    _ptr(mask), float(n_norm), _ptr(h), _ptr(tape), _ptr(loss), _ptr(gh), _ptr(gw), _ptr(gb),
                                    _stream(y0))
    lib.check(rc, "hode_fixed_fwd_sse")
    return loss, gh, gw, gb, h, tape


def fixed_bwd(lib, pb: Problem, grid, t_eval, grad_h, tape):
    D = pb.cfg.latent_dim
    grad_h = _f32c(grad_h)
    dev = grad_h.device
    gy0 = torch.empty(pb.n_traj, D, dtype=torch.float32, device=dev)
    gp = torch.empty_like(pb.params)
    with _on(grad_h):
        rc = lib.hode_fixed_bwd(C.byref(pb.cfg), pb.n_groups, pb.batch, _ptr(pb.dose_amt), _ptr(pb.dose_t),
                                pb.dose_t.stride(0), _ptr(pb.params), _ptr(pb.pset), pb.params.shape[0], _ptr(grid),
                                grid.numel(), _ptr(t_eval), t_eval.numel(), _ptr(grad_h), _ptr(tape), _ptr(gy0), _ptr(gp),
                                _stream(grad_h))
    lib.check(rc, "hode_fixed_bwd")
    return gy0, gp


def fixed_adjoint(lib, pb: Problem, adj_grid, adj_count, h, grad_h):
    """Continuous adjoint of a fixed-grid solve (``hode_fixed_adjoint``): no tape, needs the forward solution ``h``."""
    D = pb.cfg.latent_dim
    grad_h, h = _f32c(grad_h), _f32c(h)
    dev = grad_h.device
    n_t = h.shape[0]
    gy0 = torch.empty(pb.n_traj, D, dtype=torch.float32, device=dev)
    gp = torch.empty_like(pb.params)
    with _on(grad_h):
        rc = lib.hode_fixed_adjoint(C.byref(pb.cfg), pb.n_groups, pb.batch, _ptr(pb.dose_amt), _ptr(pb.dose_t),
                                    pb.dose_t.stride(0), _ptr(pb.params), _ptr(pb.pset), pb.params.shape[0],
                                    _ptr(adj_grid), adj_grid.numel(), _ptr(adj_count), n_t, _ptr(h), _ptr(grad_h),
                                    _ptr(gy0), _ptr(gp), _stream(grad_h))
    lib.check(rc, "hode_fixed_adjoint")
    return gy0, gp


def dopri5_fwd(lib, pb: Problem, y0, t_eval64, tape_capacity):
    """Returns ``h, stats [n_ctrl, 4] i32, (tape_t, tape_y) or None``."""
    D = pb.cfg.latent_dim
    y0 = _f32c(y0)
    dev = y0.device
    n_t = t_eval64.numel()
    n_ctrl = pb.n_traj if pb.cfg.controller == L.CTRL_TRAJ else pb.n_groups
    h = torch.empty(n_t, pb.n_traj, D, dtype=torch.float32, device=dev)
    stats = torch.zeros(n_ctrl, 4, dtype=torch.int32, device=dev)
    tape_t = tape_y = None
    if tape_capacity:
        tape_t = torch.empty(n_ctrl, tape_capacity, 2, dtype=torch.float64, device=dev)
        tape_y = torch.empty(tape_capacity, pb.n_traj, D, dtype=torch.float32, device=dev)
    with _on(y0):
        rc = lib.hode_dopri5_fwd(C.byref(pb.cfg), pb.n_groups, pb.batch, _ptr(y0), _ptr(pb.dose_amt), _ptr(pb.dose_t),
                                 pb.dose_t.stride(0), _ptr(pb.params), _ptr(pb.pset), _ptr(t_eval64), n_t, _ptr(h),
                                 _ptr(tape_t), _ptr(tape_y), int(tape_capacity or 0), _ptr(stats), _stream(y0))
    lib.check(rc, "hode_dopri5_fwd")
    return h, stats, (tape_t, tape_y) if tape_capacity else None


def dopri5_bwd(lib, pb: Problem, t_eval64, grad_h, tape, stats):
    D = pb.cfg.latent_dim
    grad_h = _f32c(grad_h)
    dev = grad_h.device
    tape_t, tape_y = tape
    gy0 = torch.empty(pb.n_traj, D, dtype=torch.float32, device=dev)
    gp = torch.empty_like(pb.params)
    with _on(grad_h):
        rc = lib.hode_dopri5_bwd(C.byref(pb.cfg), pb.n_groups, pb.batch, _ptr(pb.dose_amt), _ptr(pb.dose_t),
                                 pb.dose_t.stride(0), _ptr(pb.params), _ptr(pb.pset), pb.params.shape[0], _ptr(t_eval64),
                                 t_eval64.numel(), _ptr(grad_h), _ptr(tape_t), _ptr(tape_y), tape_t.shape[1], _ptr(stats),
                                 _ptr(gy0), _ptr(gp), _stream(grad_h))
    lib.check(rc, "hode_dopri5_bwd")
    return gy0, gp


def dopri5_adjoint(lib, pb: Problem, t_eval64, h, grad_h):
    """Adaptive continuous adjoint (``hode_dopri5_adjoint``): no tape, needs the forward solution ``h``.
    Returns ``grad_y0, grad_params, stats [n_ctrl, 4]``."""
    D = pb.cfg.latent_dim
    grad_h, h = _f32c(grad_h), _f32c(h)
    dev = grad_h.device
    n_ctrl = pb.n_traj if pb.cfg.controller == L.CTRL_TRAJ else pb.n_groups
    gy0 = torch.empty(pb.n_traj, D, dtype=torch.float32, device=dev)
    gp = torch.empty_like(pb.params)
    stats = torch.zeros(n_ctrl, 4, dtype=torch.int32, device=dev)
    with _on(grad_h):
        rc = lib.hode_dopri5_adjoint(C.byref(pb.cfg), pb.n_groups, pb.batch, _ptr(pb.dose_amt), _ptr(pb.dose_t),
                                     pb.dose_t.stride(0), _ptr(pb.params), _ptr(pb.pset), pb.params.shape[0], _ptr(t_eval64),
                                     t_eval64.numel(), _ptr(h), _ptr(grad_h), _ptr(gy0), _ptr(gp), _ptr(stats), _stream(grad_h))
    lib.check(rc, "hode_dopri5_adjoint")
    return gy0, gp, stats


def decode_sse(lib, h, W, b, x, mask, n_norm, want_grads=True):
    """Fused ``output_function`` + masked SSE.  Returns ``loss [1]`` and (grad_h, grad_W, grad_b) unit gradients."""
    n_t, n_traj, D = h.shape
    obs = W.shape[0]
    assert x.shape == (n_t, n_traj, obs) and mask.shape == x.shape and x.stride() == mask.stride()
    h, W, b = _f32c(h), _f32c(W), _f32c(b)
    dev = h.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    gh = gw = gb = None
    if want_grads:
        gh = torch.empty_like(h)
        gw = torch.empty(obs, D, dtype=torch.float32, device=dev)
        gb = torch.empty(obs, dtype=torch.float32, device=dev)
    with _on(h):
        rc = lib.hode_decode_sse(D, obs, n_t, n_traj, float(n_norm), _ptr(h), _ptr(W), _ptr(b), _ptr(x), _ptr(mask),
                                 x.stride(0), x.stride(1), x.stride(2), _ptr(loss), _ptr(gh), _ptr(gw), _ptr(gb), _stream(h))
    lib.check(rc, "hode_decode_sse")
    return loss, gh, gw, gb


def crps_ensemble(lib, truth, forecasts):
    """``truth [...]``, ``forecasts [..., n_mc]`` (any strides on the last axis pair after flattening) -> CRPS ``[...]``."""
    assert forecasts.shape[:-1] == truth.shape and truth.dtype == torch.float32 and forecasts.dtype == torch.float32
    n_mc = forecasts.shape[-1]
    t = truth.contiguous().reshape(-1)
    f = forecasts.reshape(-1, n_mc)
    if f.stride(0) != n_mc * f.stride(1) and not f.is_contiguous():
        f = f.contiguous()
    out = torch.empty_like(t)
    with _on(t):
        rc = lib.hode_crps_ensemble(_ptr(t), _ptr(f), t.numel(), n_mc, f.stride(0), f.stride(1), _ptr(out), _stream(t))
    lib.check(rc, "hode_crps_ensemble")
    return out.reshape(truth.shape)


def decode_crps(lib, h, W, b, x, n_mc):
    """``h [n_t, n_mc * batch, D]`` (sample-major trajectories), ``x [n_t, batch, obs]`` -> CRPS ``[n_t, batch, obs]``."""
    n_t, n_traj, D = h.shape
    obs = W.shape[0]
    assert n_traj % n_mc == 0
    batch = n_traj // n_mc
    assert x.shape == (n_t, batch, obs) and x.dtype == torch.float32
    h, W, b = _f32c(h), _f32c(W), _f32c(b)
    out = torch.empty(n_t, batch, obs, dtype=torch.float32, device=h.device)
    with _on(h):
        rc = lib.hode_decode_crps(D, obs, n_t, batch, n_mc, _ptr(h), _ptr(W), _ptr(b), _ptr(x), x.stride(0), x.stride(1),
                                  x.stride(2), _ptr(out), _stream(h))
    lib.check(rc, "hode_decode_crps")
    return out


# ---- real-data fields (hode_real_*) ----------------------------------------------------------------------------------
def real_dose_tables(lib, field, action, params):
    """``action [T, B, 1]`` -> dose table ``[2 or 1, T + 1, B]`` (``params`` supplies ``kel`` for RocheODEReal)."""
    assert action.dim() == 3 and action.shape[2] == 1 and action.dtype == torch.float32
    T, B = action.shape[0], action.shape[1]
    tab = torch.empty(2 if field == L.FIELD_ROCHE_REAL else 1, T + 1, B, dtype=torch.float32, device=action.device)
    with _on(action):
        rc = lib.hode_real_dose_tables(int(field), _ptr(action), action.stride(0), action.stride(1), T, B, _ptr(params),
                                       _ptr(tab), _stream(action))
    lib.check(rc, "hode_real_dose_tables")
    return tab


def real_fixed_fwd(lib, field, D, hidden, method, perturb, y0, tab, params, grid, t_eval, want_tape):
    y0 = _f32c(y0)
    B = y0.shape[0]
    n_t, n_grid = t_eval.numel(), grid.numel()
    h = torch.empty(n_t, B, D, dtype=torch.float32, device=y0.device)
    tape = torch.empty(max(n_grid - 1, 0), B, D, dtype=torch.float32, device=y0.device) if want_tape else None
    with _on(y0):
        rc = lib.hode_real_fixed_fwd(int(field), D, int(hidden), int(method), int(bool(perturb)), B, _ptr(y0), _ptr(tab),
                                     tab.shape[1] - 1, _ptr(params), _ptr(grid), n_grid, _ptr(t_eval), n_t, _ptr(h),
                                     _ptr(tape), _stream(y0))
    lib.check(rc, "hode_real_fixed_fwd")
    return h, tape


def real_fixed_bwd(lib, field, D, hidden, method, perturb, tab, params, grid, t_eval, grad_h, tape):
    grad_h = _f32c(grad_h)
    B = grad_h.shape[1]
    gy0 = torch.empty(B, D, dtype=torch.float32, device=grad_h.device)
    gp = torch.empty_like(params)
    with _on(grad_h):
        rc = lib.hode_real_fixed_bwd(int(field), D, int(hidden), int(method), int(bool(perturb)), B, _ptr(tab),
                                     tab.shape[1] - 1, _ptr(params), _ptr(grid), grid.numel(), _ptr(t_eval), t_eval.numel(),
                                     _ptr(grad_h), _ptr(tape), _ptr(gy0), _ptr(gp), _stream(grad_h))
    lib.check(rc, "hode_real_fixed_bwd")
    return gy0, gp
