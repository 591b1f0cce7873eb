"""Drop-in model classes of the hot path: same names, constructor signatures, attributes and ``state_dict`` keys as
``/root/reference/model.py`` (``RocheODE`` 446-555, ``NeuralODE`` 969-1026, ``RocheExpertDecoder`` 1030-1121), with
the solve routed to the fused sm_100a kernels through :func:`hybrid_ode_neurips_2021_b200.solver.odeint`.

What differs from the reference, deliberately:
* ``set_action`` is one kernel launch (plus one small host read to check that every patient has the same number of
  doses, which the reference enforces through ``torch.stack``) instead of an O(B) Python loop;
* ``device=None`` means the current CUDA device (the reference hard-codes ``cuda:1``, ``global_config.py:7``);
* ``RocheExpertDecoder`` takes an optional ``solver_options`` dict that is forwarded to ``odeint`` (the reference
  never forwards a step size: its ``ode_step_size`` only lands in a dead ``options['h']``, ``model.py:1076`` -- kept);
* ``forward(t, y)`` of the fields stays an eager PyTorch function for callers that evaluate the field directly; it is
  not what ``odeint`` executes.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .solver import odeint, odeint_adjoint, odeint_sse

DTYPE = torch.float32


class RochConfig(NamedTuple):
    """Expert-ODE defaults (``sim_config.py:4-18``)."""

    HillCure: float = 2
    HillPatho: float = 2
    ec50_patho: float = 1
    emax_patho: float = 1
    k_dexa: float = 1
    k_discure_immunereact: float = 1
    k_discure_immunity: float = 1
    k_disprog: float = 1
    k_immune_disease: float = 1
    k_immune_feedback: float = 1
    k_immune_off: float = 1
    k_immunity: float = 1
    kel: float = 1


def _default_device():
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


class _DoseMixin:
    """Vectorised ``set_action``: ``dosage [B]``, ``times [B, n_dose]`` (same attributes as the reference)."""

    def set_action(self, action):
        if action.shape[1] == 0:  # the reference's torch.stack over an empty list of patients (model.py:507)
            raise RuntimeError("stack expects a non-empty TensorList")
        if action.is_cuda:
            amt, idx, cnt = ops.dose_schedule(L.get_lib(), action if action.dtype == torch.float32 else action.float())
            lo, hi = (int(v) for v in torch.stack(torch.aminmax(cnt)).tolist())
            if lo != hi:
                raise RuntimeError("stack expects each tensor to be equal size, but patients have between {} and {} "
                                   "doses".format(lo, hi))
            idx = idx[:, :hi]
            self.dosage = amt.to(action.dtype)
            self.times = idx.to(torch.int64) * self.step_size
            self._dose_t_f32 = self.times.to(torch.float32).contiguous()
        else:  # host tensors: the reference's own loop (not a compute path of this package)
            self.dosage = torch.max(action[..., 0], dim=0)[0]
            rows = [torch.where(action[:, i, 0] != 0)[0] * self.step_size for i in range(action.shape[1])]
            self.times = torch.stack(rows, dim=0)
            self._dose_t_f32 = None


class RocheODE(_DoseMixin, nn.Module):
    def __init__(self, latent_dim, action_dim, t_max, step_size, ablate=False, device=None, dtype=DTYPE):
        super().__init__()
        assert action_dim == 1
        self.action_dim = action_dim
        self.latent_dim = int(latent_dim)
        self.expert_dim = 4
        self.ml_dim = self.latent_dim - self.expert_dim
        self.expanded = self.ml_dim > 0
        self.ablate = ablate
        self.device = _default_device() if device is None else device
        self.t_max = t_max
        self.step_size = step_size
        dc = RochConfig()
        for name in dc._fields:  # registration order == the reference's named_parameters() order
            setattr(self, name, nn.Parameter(torch.tensor(getattr(dc, name), device=self.device, dtype=dtype)))
        if self.ablate:
            self.theta_1 = nn.Parameter(torch.tensor(1, device=self.device, dtype=dtype))
            self.theta_2 = nn.Parameter(torch.tensor(2, device=self.device, dtype=dtype))
        if self.expanded:
            self.ml_net = nn.Sequential(nn.Linear(self.latent_dim, self.ml_dim), nn.Tanh()).to(self.device)
        else:
            self.ml_net = nn.Identity().to(self.device)
        self.times = None
        self.dosage = None
        self._dose_t_f32 = None

    def dose_at_time(self, t):
        on = t >= self.times
        return self.dosage * torch.sum(torch.exp(self.kel * (self.times - t) * on) * on, dim=-1)

    def forward(self, t, y):
        dis, react, imm, dose2 = y[:, 0], y[:, 1], y[:, 2], y[:, 3]
        if not self.ablate:
            dose = self.dose_at_time(t)
            rp = react ** self.HillPatho
            d1 = dis * self.k_disprog - dis * imm ** self.HillCure * self.k_discure_immunity \
                - dis * react * self.k_discure_immunereact
            d2 = dis * self.k_immune_disease - react * self.k_immune_off + dis * react * self.k_immune_feedback \
                + (rp * self.emax_patho) / (self.ec50_patho ** self.HillPatho + rp) - dose2 * react * self.k_dexa
            d3 = react * self.k_immunity
            d4 = self.kel * dose - self.kel * dose2
        else:
            d1, d2, d3, d4 = react, -1.0 * dis * self.theta_1, dose2, -1.0 * imm * self.theta_2
        cols = [d1[..., None], d2[..., None], d3[..., None], d4[..., None]]
        if self.expanded:
            cols.append(self.ml_net(y))
        return torch.cat(cols, dim=-1)


class NeuralODE(_DoseMixin, nn.Module):
    def __init__(self, latent_dim, action_dim, t_max, step_size, device=None, dtype=DTYPE):
        super().__init__()
        assert action_dim == 1
        self.action_dim = action_dim
        self.latent_dim = int(latent_dim)
        self.expert_dim = 4
        self.ml_dim = self.latent_dim
        self.device = _default_device() if device is None else device
        self.t_max = t_max
        self.step_size = step_size
        self.kel = nn.Parameter(torch.tensor(RochConfig().kel, device=self.device, dtype=dtype))
        d = self.latent_dim
        self.ml_net = nn.Sequential(nn.Linear(d + 1, d * 10), nn.Tanh(), nn.Linear(d * 10, d), nn.Tanh()).to(self.device)
        self.times = None
        self.dosage = None
        self._dose_t_f32 = None

    def dose_at_time(self, t):
        return self.dosage * torch.sum(self.times == t, dim=-1)

    def forward(self, t, y):
        return self.ml_net(torch.cat([y, self.dose_at_time(t)[:, None]], dim=-1))


class RocheExpertDecoder(nn.Module):
    def __init__(self, obs_dim, latent_dim, action_dim, t_max, step_size, roche=True, ablate=False, method="dopri5",
                 ode_step_size=None, device=None, dtype=DTYPE, solver_options=None, adjoint=False, adjoint_options=None):
        super().__init__()
        self.adjoint = bool(adjoint)  # True: backward by the continuous adjoint (model.py:9's alternative import)
        self.adjoint_options = adjoint_options  # e.g. {'norm': 'seminorm'} for the adaptive (dopri5) adjoint
        self.time_dim = int(t_max / step_size)
        self.obs_dim = obs_dim
        self.latent_dim = latent_dim
        self.action_dim = action_dim
        self.t_max = t_max
        self.step_size = step_size
        self.roche = roche
        self.ablate = ablate
        if roche:
            self.model_name = "ExpertDecoder" if latent_dim == 4 else "HybridDecoder"
        else:
            self.model_name = "NeuralODEDecoder"
        if self.ablate:
            self.model_name = self.model_name + "Ablate"
            print("Running ablation study")
        self.device = _default_device() if device is None else device
        self.t = torch.arange(0, t_max + step_size, step_size, device=self.device, dtype=dtype)
        # same keys as the reference's (mostly dead) option dict; only method / rtol / atol reach the solver
        self.options = {
            "method": method, "h": ode_step_size, "t0": 0.0, "t1": t_max + step_size, "rtol": 1e-7, "atol": 1e-8,
            "print_neval": True, "neval_max": 1000000, "safety": None, "t_eval": self.t,
            "interpolation_method": "cubic", "regenerate_graph": False,
        }
        self.solver_options = solver_options
        self.output_function = nn.Sequential(nn.Linear(self.latent_dim, self.obs_dim, bias=True)).to(self.device)
        if roche:
            self.ode = RocheODE(latent_dim, action_dim, t_max, step_size, ablate=self.ablate, device=self.device)
        else:
            self.ode = NeuralODE(latent_dim, action_dim, t_max, step_size, self.device)

    def solve(self, init, a):
        """Latent trajectories ``h [T, B, D]`` only (no read-out)."""
        self.ode.set_action(a)
        kw = dict(rtol=self.options["rtol"], atol=self.options["atol"], method=self.options["method"], options=self.solver_options)
        if self.adjoint:
            return odeint_adjoint(self.ode, init, self.t, adjoint_options=self.adjoint_options, **kw)
        return odeint(self.ode, init, self.t, **kw)

    def forward(self, init, a):
        h = self.solve(init, a)
        return self.output_function(h), h

    def loss(self, init, a, x, mask, n_norm=None):
        """The likelihood of ``VariationalInference.loss`` (``model.py:1172-1179``): ``forward`` followed by
        ``sum((x - x_hat)**2 * mask) / x.shape[1]``, as one fused launch where the kernel exists (:func:`solver.odeint_sse`);
        ``x_hat`` and ``h`` are not materialised.  The continuous-adjoint decoder keeps its two-call path."""
        if self.adjoint:
            from .loss import masked_sse

            return masked_sse(self, self.solve(init, a), x, mask, n_norm)
        self.ode.set_action(a)
        lin = self.output_function[0]
        return odeint_sse(self.ode, init, self.t, lin.weight, lin.bias, x, mask, n_norm=n_norm, rtol=self.options["rtol"],
                          atol=self.options["atol"], method=self.options["method"], options=self.solver_options)
