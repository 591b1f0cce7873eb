"""CPU oracle: the reference's vector fields, decoder and masked-SSE loss restated as plain PyTorch functions.

TEST INFRASTRUCTURE ONLY (see ``oracle/odeint.py``).  ``/root/reference`` does not exist on the GPU box, so the
checker cannot import ``model.py`` there; these restatements travel instead.  Each one follows the reference
operation-for-operation (same association order, same dtype promotions), and ``tests/test_oracle_fields.py`` checks
them BIT-FOR-BIT against the reference's own modules whenever ``/root/reference`` is importable, and against the
committed fixtures in ``tests/golden/`` everywhere.

Parameter containers are ordinary ``nn.Module``s with the reference's ``state_dict`` keys so that weights can be
exchanged with both the reference classes and the CUDA drop-in.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import odeint as _oi

# named_parameters() order of the reference's RocheODE (model.py:468-481); RochConfig defaults (sim_config.py:4-18)
EXPERT_NAMES = (
    "HillCure",
    "HillPatho",
    "ec50_patho",
    "emax_patho",
    "k_dexa",
    "k_discure_immunereact",
    "k_discure_immunity",
    "k_disprog",
    "k_immune_disease",
    "k_immune_feedback",
    "k_immune_off",
    "k_immunity",
    "kel",
)
EXPERT_DEFAULTS = (2.0, 2.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)


def dose_schedule(action: torch.Tensor, step_size):
    """``set_action`` (model.py:495-507, 1001-1013): one amount per patient, one row of dose times per patient.

    ``action`` is ``[T, B, 1]``.  Returns ``dosage [B]`` (``action.dtype``) and ``times [B, n_dose]`` (int64 when
    ``step_size`` is an int).  Patients must all have the same number of non-zero actions (``torch.stack``).
    """
    a = action[..., 0]
    dosage = torch.max(a, dim=0)[0]
    rows = []
    for b in range(a.shape[1]):
        rows.append(torch.where(a[:, b] != 0)[0] * step_size)
    return dosage, torch.stack(rows, dim=0)


class OracleRocheODE(nn.Module):
    """Expert PK/PD field + optional latent MLP (model.py:446-555); ``ablate=True`` is the ablation study's expert part
    (model.py:545-549)."""

    def __init__(self, latent_dim, step_size=1, dtype=torch.float32, ablate=False):
        super().__init__()
        self.latent_dim = int(latent_dim)
        self.ml_dim = self.latent_dim - 4
        self.step_size = step_size
        self.ablate = bool(ablate)
        for name, val in zip(EXPERT_NAMES, EXPERT_DEFAULTS):
            setattr(self, name, nn.Parameter(torch.tensor(val, dtype=dtype)))
        if self.ablate:
            self.theta_1 = nn.Parameter(torch.tensor(1.0, dtype=dtype))
            self.theta_2 = nn.Parameter(torch.tensor(2.0, dtype=dtype))
        if self.ml_dim > 0:
            self.ml_net = nn.Sequential(nn.Linear(self.latent_dim, self.ml_dim), nn.Tanh()).to(dtype)
        else:
            self.ml_net = nn.Identity()
        self.times = None
        self.dosage = None

    def set_action(self, action):
        self.dosage, self.times = dose_schedule(action, self.step_size)

    def dose_at_time(self, t):
        on = t >= self.times
        return self.dosage * torch.sum(torch.exp(self.kel * (self.times - t) * on) * on, dim=-1)

    def forward(self, t, y):
        dis, react, imm, dose2 = y[:, 0], y[:, 1], y[:, 2], y[:, 3]
        if self.ablate:
            d1, d2, d3, d4 = react, -1.0 * dis * self.theta_1, dose2, -1.0 * imm * self.theta_2
            if self.ml_dim > 0:
                return torch.cat([d1[..., None], d2[..., None], d3[..., None], d4[..., None], self.ml_net(y)], dim=-1)
            return torch.stack([d1, d2, d3, d4], dim=-1)
        dose = self.dose_at_time(t)
        d1 = (
            dis * self.k_disprog
            - dis * imm ** self.HillCure * self.k_discure_immunity
            - dis * react * self.k_discure_immunereact
        )
        d2 = (
            dis * self.k_immune_disease
            - react * self.k_immune_off
            + dis * react * self.k_immune_feedback
            + (react ** self.HillPatho * self.emax_patho) / (self.ec50_patho ** self.HillPatho + react ** self.HillPatho)
            - dose2 * react * self.k_dexa
        )
        d3 = react * self.k_immunity
        d4 = self.kel * dose - self.kel * dose2
        if self.ml_dim > 0:
            return torch.cat([d1[..., None], d2[..., None], d3[..., None], d4[..., None], self.ml_net(y)], dim=-1)
        return torch.stack([d1, d2, d3, d4], dim=-1)


class OracleNeuralODE(nn.Module):
    """Pure neural field on ``[y, Dose]`` with an impulse dose (model.py:969-1026)."""

    def __init__(self, latent_dim, step_size=1, dtype=torch.float32):
        super().__init__()
        self.latent_dim = int(latent_dim)
        self.step_size = step_size
        self.kel = nn.Parameter(torch.tensor(1.0, dtype=dtype))  # unused by forward, but a state_dict key
        d = self.latent_dim
        self.ml_net = nn.Sequential(nn.Linear(d + 1, d * 10), nn.Tanh(), nn.Linear(d * 10, d), nn.Tanh()).to(dtype)
        self.times = None
        self.dosage = None

    def set_action(self, action):
        self.dosage, self.times = dose_schedule(action, self.step_size)

    def dose_at_time(self, t):
        return self.dosage * torch.sum(self.times == t, dim=-1)

    def forward(self, t, y):
        return self.ml_net(torch.cat([y, self.dose_at_time(t)[:, None]], dim=-1))


class OracleDecoder(nn.Module):
    """``RocheExpertDecoder`` (model.py:1030-1121) with an explicit ``options`` pass-through for fixed-step sweeps."""

    def __init__(self, obs_dim, latent_dim, t_max=14, step_size=1, roche=True, method="dopri5", options=None,
                 rtol=1e-7, atol=1e-8, dtype=torch.float32):
        super().__init__()
        self.t = torch.arange(0, t_max + step_size, step_size, dtype=dtype)
        self.method, self.options, self.rtol, self.atol = method, options, rtol, atol
        self.output_function = nn.Sequential(nn.Linear(latent_dim, obs_dim, bias=True)).to(dtype)
        self.ode = OracleRocheODE(latent_dim, step_size, dtype) if roche else OracleNeuralODE(latent_dim, step_size, dtype)

    def forward(self, init, a, trace=None):
        self.ode.set_action(a)
        opts = dict(self.options or {})
        if trace is not None:
            opts["trace"] = trace
        h = _oi.odeint(self.ode, init, self.t, rtol=self.rtol, atol=self.atol, method=self.method, options=opts)
        return self.output_function(h), h


def masked_sse(x, x_hat, mask):
    """Likelihood term of ``VariationalInference.loss`` (model.py:1179): ``sum((x - x_hat)^2 * mask) / B``."""
    return torch.sum((x - x_hat) ** 2 * mask) / x.shape[1]


# ----------------------------------------------------------------------------------------------------------------------
# real-data (ICU) fields: model.py:570-769, decoder model.py:772-862
# ----------------------------------------------------------------------------------------------------------------------
class OracleRocheODEReal(nn.Module):
    """``RocheODEReal`` (model.py:570-657): learned expert nets + GRU-ODE latent block; dose = sum over ALL past doses."""

    def __init__(self, latent_dim, hidden_dim, dtype=torch.float32):
        super().__init__()
        self.latent_dim, self.hidden_dim, self.expert_dim = int(latent_dim), int(hidden_dim), 4
        H = self.hidden_dim
        self.dx1_net = nn.Sequential(nn.Linear(3, H), nn.Tanh(), nn.Linear(H, 1), nn.Tanh()).to(dtype)
        self.dx2_net = nn.Sequential(nn.Linear(2, H), nn.Tanh(), nn.Linear(H, 1), nn.Tanh()).to(dtype)
        self.expert_only = self.latent_dim == self.expert_dim
        if not self.expert_only:
            m = self.latent_dim - self.expert_dim
            self.lin_hh = nn.Linear(m, m, bias=False).to(dtype)
            self.lin_hz = nn.Linear(m, m, bias=False).to(dtype)
            self.lin_hr = nn.Linear(m, m, bias=False).to(dtype)
        self.k_immunity = nn.Parameter(torch.tensor(1.0, dtype=dtype))
        self.kel = nn.Parameter(torch.tensor(0.2, dtype=dtype))
        self.kel2 = nn.Parameter(torch.tensor(0.2, dtype=dtype))
        self.dosage = None
        self.times = None

    def set_action_static(self, action, static):
        self.dosage = action
        self.times = torch.cumsum(torch.ones_like(action), dim=0)

    def dose_at_time(self, t):
        inside_exp = self.kel * (self.times - t) * (t >= self.times)
        return torch.sum(self.dosage * torch.exp(inside_exp) * (t >= self.times), dim=(0, 2))

    def forward(self, t, y):
        react, dose2 = y[:, 1], y[:, 3]
        dose = self.dose_at_time(t)
        d1 = self.dx1_net(y[:, :3])
        d2 = self.dx2_net(y[:, :2])
        d3 = (react * self.k_immunity)[..., None]
        d4 = (self.kel * dose - self.kel2 * dose2)[..., None]
        if self.expert_only:
            return torch.cat([d1, d2, d3, d4], dim=-1)
        x = 0
        h = y[..., self.expert_dim:]
        r = torch.sigmoid(x + self.lin_hr(h))
        z = torch.sigmoid(x + self.lin_hz(h))
        u = torch.tanh(x + self.lin_hh(r * h))
        return torch.cat([d1, d2, d3, d4, (1 - z) * (u - h)], dim=-1)


class OracleNeuralODEReal(nn.Module):
    """``NeuralODEReal`` (model.py:717-769) and, with ``second=True``, ``NeuralODEReal2nd`` (model.py:660-714)."""

    def __init__(self, latent_dim, hidden_dim, second=False, dtype=torch.float32):
        super().__init__()
        self.latent_dim, self.hidden_dim, self.second = int(latent_dim), int(hidden_dim), bool(second)
        out = self.latent_dim // 2 if second else self.latent_dim
        self.ml_net = nn.Sequential(nn.Linear(self.latent_dim + 1, self.hidden_dim), nn.Tanh(),
                                    nn.Linear(self.hidden_dim, out), nn.Tanh()).to(dtype)
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
        if self.second:
            return torch.cat([out, y[..., : (self.latent_dim // 2)]], dim=-1)
        return out


class OracleDecoderReal(nn.Module):
    """``DecoderReal`` (model.py:772-862), 2-D ``init`` branch: ``t = arange(t0 - 1, t_max)``, fixed-grid solver with
    ``options = {step_t, step_size, perturb: True}``, read-out MLP ``Linear -> ELU -> Linear``, ``x_hat = out[1:]``."""

    def __init__(self, obs_dim, latent_dim, hidden_dim, t_max, t0=24, method="midpoint", ode_step_size=1.0,
                 ode_type="hybrid", dtype=torch.float32):
        super().__init__()
        self.t_max, self.t0, self.method = t_max, t0, method
        self.output_function = nn.Sequential(nn.Linear(latent_dim, latent_dim + 1), nn.ELU(),
                                             nn.Linear(latent_dim + 1, obs_dim)).to(dtype)
        if ode_type == "neural":
            self.ode = OracleNeuralODEReal(latent_dim, hidden_dim, False, dtype)
        elif ode_type == "2nd":
            self.ode = OracleNeuralODEReal(latent_dim, hidden_dim, True, dtype)
        else:
            self.ode = OracleRocheODEReal(latent_dim, hidden_dim, dtype)
        self.t = torch.arange(t0 - 1, t_max, 1.0, dtype=dtype)
        self.options = {"step_t": self.t, "step_size": ode_step_size, "perturb": True}

    def forward(self, init, a, s):
        self.ode.set_action_static(a, s)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # fixed-grid solvers warn about the unused step_t, like torchdiffeq
            h = _oi.odeint(self.ode, init, self.t, method=self.method, options=self.options, rtol=1e-7, atol=1e-8)
        return self.output_function(h)[1:], h
