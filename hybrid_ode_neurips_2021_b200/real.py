"""Real-data (ICU cohort) drop-ins: ``RocheODEReal`` / ``NeuralODEReal`` / ``NeuralODEReal2nd`` / ``DecoderReal`` with the
constructor signatures, attributes and ``state_dict`` keys of ``/root/reference/model.py:570-862``, integrated by the
fused fixed-grid kernels of ``csrc/hode_real.cu``.

The reference evaluates ``dose_at_time`` as an O(T) sum over every hourly dose for each vector-field call
(``model.py:653-657``) and calls ``int(t)`` on a device scalar (``:696, 753`` -- a host sync per call); here a per-launch
dose table makes the input O(1) and the whole solve is one launch.  ``DecoderReal`` only ever integrates these fields
with ``euler`` / ``midpoint`` / ``rk4`` (``experiments/real.sh:9-17``); ``dopri5`` on them raises ``NotImplementedError``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import ops

DTYPE = torch.float32


def _default_device():
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


# ----------------------------------------------------------------------------------------------------------------------
def real_field_kind(func):
    """hode_field of a real-data vector field (ours or the reference's own class, recognised by structure) or None."""
    if not isinstance(func, nn.Module):
        return None
    if hasattr(func, "dx1_net") and hasattr(func, "dx2_net") and hasattr(func, "kel2"):
        return L.FIELD_ROCHE_REAL
    if hasattr(func, "set_action_static") and hasattr(func, "ml_net") and isinstance(func.ml_net, nn.Sequential) \
            and len(func.ml_net) == 4 and hasattr(func, "static_dim"):
        out = func.ml_net[2].out_features
        if out == int(func.latent_dim):
            return L.FIELD_NEURAL_REAL
        if out == int(func.latent_dim) // 2:
            return L.FIELD_NEURAL_REAL_2ND
    return None


def pack_real_params_list(func, kind):
    """The parameters in packed order (layout of include/hode.h == ``state_dict`` order of the reference classes)."""
    if kind == L.FIELD_ROCHE_REAL:
        parts = [func.k_immunity, func.kel, func.kel2]
        for net in (func.dx1_net, func.dx2_net):
            parts += [net[0].weight, net[0].bias, net[2].weight, net[2].bias]
        if int(func.latent_dim) > 4:
            parts += [func.lin_hh.weight, func.lin_hz.weight, func.lin_hr.weight]
        return parts
    l1, l2 = func.ml_net[0], func.ml_net[2]
    return [l1.weight, l1.bias, l2.weight, l2.bias]


def pack_real_params(func, kind) -> torch.Tensor:
    """Differentiable packed parameter vector."""
    return torch.cat([p.reshape(-1) for p in pack_real_params_list(func, kind)]).to(torch.float32)


def real_action(func, kind):
    a = func.dosage if kind == L.FIELD_ROCHE_REAL else func.action
    if a is None:
        raise RuntimeError("set_action_static must be called before integrating (model.py:834)")
    if a.dim() != 3 or a.shape[2] != 1:
        raise NotImplementedError("real-data fields take one action channel (run_real.py: action_dim = 1)")
    return a


class _RealFixedSolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, packed, meta, grid, t_eval, need_grad):
        lib = L.get_lib()
        kind, D, H, method, perturb, action = meta
        params = packed.detach().contiguous()
        tab = ops.real_dose_tables(lib, kind, action.detach().float(), params)
        h, tape = ops.real_fixed_fwd(lib, kind, D, H, method, perturb, y0.detach(), tab, params, grid, t_eval, need_grad)
        ctx.meta, ctx.tab, ctx.params, ctx.grid, ctx.t_eval, ctx.tape = meta, tab, params, grid, t_eval, tape
        return h

    @staticmethod
    def backward(ctx, grad_h):
        if ctx.tape is None:
            raise RuntimeError("backward through a solve that was run without a tape")
        kind, D, H, method, perturb, _ = ctx.meta
        gy0, gp = ops.real_fixed_bwd(L.get_lib(), kind, D, H, method, perturb, ctx.tab, ctx.params, ctx.grid, ctx.t_eval,
                                     grad_h, ctx.tape)
        return gy0, gp, None, None, None, None


def solve_real(func, kind, y0, method, perturb, grid, t_eval, need_grad):
    D = int(func.latent_dim)
    H = int(func.hidden_dim)
    lib = L.get_lib()
    if int(lib.hode_real_param_count(int(kind), D, H)) < 0:
        raise NotImplementedError("real-data field with latent_dim {} / hidden_dim {} has no compiled kernel "
                                  "(latent 4 / 20, 2nd-order 8 / 40, hidden <= 64)".format(D, H))
    action = real_action(func, kind)
    if action.shape[1] != y0.shape[0]:
        raise RuntimeError("actions cover {} patients but y0 has {}".format(action.shape[1], y0.shape[0]))
    packed = pack_real_params(func, kind)
    meta = (kind, D, H, L.METHODS[method], perturb, action.to(y0.device))
    return _RealFixedSolve.apply(y0, packed, meta, grid, t_eval, need_grad)


# ----------------------------------------------------------------------------------------------------------------------
class RocheODEReal(nn.Module):
    def __init__(self, latent_dim, action_dim, static_dim, hidden_dim, t_max, step_size, device=None, dtype=DTYPE):
        super().__init__()
        self.action_dim, self.latent_dim = int(action_dim), int(latent_dim)
        self.static_dim, self.hidden_dim = int(static_dim), int(hidden_dim)
        self.dosage = None
        self.times = None
        self.device = _default_device() if device is None else device
        self.t_max, self.step_size = t_max, step_size
        H = self.hidden_dim
        self.dx1_net = nn.Sequential(nn.Linear(3, H), nn.Tanh(), nn.Linear(H, 1), nn.Tanh()).to(self.device)
        self.dx2_net = nn.Sequential(nn.Linear(2, H), nn.Tanh(), nn.Linear(H, 1), nn.Tanh()).to(self.device)
        self.expert_dim = 4
        self.expert_only = self.latent_dim == self.expert_dim
        if not self.expert_only:
            m = self.latent_dim - self.expert_dim
            self.lin_hh = nn.Linear(m, m, bias=False).to(self.device)
            self.lin_hz = nn.Linear(m, m, bias=False).to(self.device)
            self.lin_hr = nn.Linear(m, m, bias=False).to(self.device)
        self.k_immunity = nn.Parameter(torch.tensor(1, device=self.device, dtype=dtype))
        self.kel = nn.Parameter(torch.tensor(0.2, device=self.device, dtype=dtype))
        self.kel2 = nn.Parameter(torch.tensor(0.2, device=self.device, dtype=dtype))

    def set_action_static(self, action, static):
        self.dosage = action
        self.times = torch.cumsum(torch.ones_like(action), dim=0)

    def dose_at_time(self, t):
        inside_exp = self.kel * (self.times - t) * (t >= self.times)
        return torch.sum(self.dosage * torch.exp(inside_exp) * (t >= self.times), dim=(0, 2))

    def forward(self, t, y):  # eager evaluation for callers that use the field directly; odeint runs the fused kernels
        d1, d2 = self.dx1_net(y[:, :3]), self.dx2_net(y[:, :2])
        d3 = (y[:, 1] * self.k_immunity)[..., None]
        d4 = (self.kel * self.dose_at_time(t) - self.kel2 * y[:, 3])[..., None]
        if self.expert_only:
            return torch.cat([d1, d2, d3, d4], dim=-1)
        h = y[..., self.expert_dim:]
        r, z = torch.sigmoid(self.lin_hr(h)), torch.sigmoid(self.lin_hz(h))
        u = torch.tanh(self.lin_hh(r * h))
        return torch.cat([d1, d2, d3, d4, (1 - z) * (u - h)], dim=-1)


class _NeuralRealBase(nn.Module):
    SECOND = False

    def __init__(self, latent_dim, action_dim, static_dim, hidden_dim, t_max, step_size, device=None, dtype=DTYPE):
        super().__init__()
        self.action_dim, self.latent_dim = int(action_dim), int(latent_dim)
        self.static_dim, self.hidden_dim = int(static_dim), int(hidden_dim)
        self.device = _default_device() if device is None else device
        self.t_max, self.step_size = t_max, step_size
        out = self.latent_dim // 2 if self.SECOND else self.latent_dim
        self.ml_net = nn.Sequential(nn.Linear(self.latent_dim + self.action_dim, self.hidden_dim), nn.Tanh(),
                                    nn.Linear(self.hidden_dim, out), nn.Tanh()).to(self.device)
        self.action = None
        self.static = None

    def set_action_static(self, action, static):
        self.action = action
        self.static = static[0, :, :]

    def dose_at_time(self, t):
        t_int = int(t)
        if t_int >= self.action.shape[0]:
            return torch.zeros_like(self.action[0, :, :])
        return torch.cumsum(self.action, dim=0)[t_int, :, :]

    def forward(self, t, y):
        out = self.ml_net(torch.cat([y, self.dose_at_time(t)], dim=-1))
        if self.SECOND:
            return torch.cat([out, y[..., : (self.latent_dim // 2)]], dim=-1)
        return out


class NeuralODEReal(_NeuralRealBase):
    SECOND = False


class NeuralODEReal2nd(_NeuralRealBase):
    SECOND = True


class DecoderReal(nn.Module):
    def __init__(self, obs_dim, latent_dim, action_dim, static_dim, hidden_dim, t_max, step_size, t0=0, method="dopri5",
                 ode_step_size=None, ode_type="neural", device=None, dtype=DTYPE):
        super().__init__()
        from .solver import odeint  # late import: solver imports this module

        self._odeint = odeint
        self.time_dim = int(t_max / step_size)
        self.obs_dim, self.latent_dim, self.action_dim = obs_dim, latent_dim, action_dim
        self.t_max, self.t0 = t_max, t0
        self.static_dim, self.hidden_dim = int(static_dim), int(hidden_dim)
        self.model_name = "DecoderReal_" + ode_type
        self.device = _default_device() if device is None else device
        self.output_function = nn.Sequential(
            nn.Linear(self.latent_dim, self.latent_dim + 1, bias=True), nn.ELU(),
            nn.Linear(self.latent_dim + 1, self.obs_dim, bias=True),
        ).to(self.device)
        cls = {"neural": NeuralODEReal, "2nd": NeuralODEReal2nd}.get(ode_type, RocheODEReal)
        self.ode = cls(latent_dim, action_dim, static_dim, hidden_dim, t_max, step_size, self.device)
        self.t = torch.arange(t0 - 1, t_max, step_size, device=self.device, dtype=dtype)
        self.rtol, self.atol = 1e-7, 1e-8
        self.options = {"step_t": self.t, "step_size": ode_step_size, "perturb": True}
        self.method = method
        self.step_size = ode_step_size

    def _dto(self, y0, t):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # "Unexpected arguments {'step_t': ...}", exactly like torchdiffeq
            return self._odeint(self.ode, y0, t, method=self.method, options=self.options, rtol=self.rtol, atol=self.atol)

    def forward(self, init, a, s):
        self.ode.set_action_static(a, s)
        if len(init.shape) == 2:
            h = self._dto(init, self.t)
        else:  # one-step-ahead mode (model.py:840-855): a fresh initial state for every interval
            h_list = [self._dto(init[i], self.t[i:(i + 2)])[-1, ...] for i in range(self.t_max - 1)]
            h = torch.stack([torch.zeros_like(h_list[0])] + h_list, dim=0)
        x_hat = self.output_function(h)[1:]
        if len(init.shape) != 2:
            x_hat[0] = 0.0
        return x_hat, h
