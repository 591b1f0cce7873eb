"""Synthetic-cohort generator with the reference's ``DataGeneratorRoche`` interface (``/root/reference/dataloader.py:10-342``),
the data format on the input side of the hot path.

The reference integrates the TRUE hybrid ODE (expert PK/PD terms + ``tanh(y @ ml_coef)`` latents) one patient at a time
with scipy's ``lsoda`` (``dataloader.py:95-164``, ~22 ms per patient: a 1 M-patient cohort would take six hours).  Here
the whole cohort is ONE launch of the fused dopri5 kernel with a per-trajectory step-size controller
(``hode_dopri5_fwd``, ``HODE_CTRL_TRAJ``): the generator's ODE is exactly the ``RocheODE`` field with
``ml_net[0].weight = ml_coef.T`` and zero bias, so no new kernel is needed.  Read-out, noise, normalisation and masking
follow ``dataloader.py:166-266`` on the device.

Random streams.  ``exact_rng=True`` (default) consumes numpy's and torch's global generators in the reference's order
(coefficients, initial conditions, one ``np.random.choice`` per patient, amounts, one ``randn(obs, T)`` per patient,
``torch.rand_like`` on the CPU), so with the same seeds every random quantity is bit-identical to the reference's and
only the latents differ (float32 dopri5 vs float64 lsoda, ~1e-4 absolute).  ``exact_rng=False`` draws the per-patient
quantities on the device instead (same distributions, different stream) for cohorts where a host loop over patients
is the bottleneck.

There is no CPU fallback: ``generate_data`` raises unless the cohort device is CUDA (tests inject the host emulation
through ``lib=``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import ops
from .model import DTYPE, _default_device

_KEYS = ("measurements", "actions", "latents", "masks")


class DataGeneratorRoche:
    def __init__(self, n_sample, obs_dim, t_max, step_size, roche_config, output_sigma, dose_max=0, latent_dim=4,
                 sparsity=0.5, output_sparsity=0.0, val_size=100, test_size=200, p_remove=0, device=None, dtype=DTYPE,
                 exact_rng=True, rtol=1e-7, atol=1e-8, lib=None):
        self.device = _default_device() if device is None else device
        self.dtype = dtype
        self.n_sample, self.obs_dim, self.latent_dim = n_sample, obs_dim, int(latent_dim)
        self.expert_dim, self.action_dim = 4, 1
        self.ml_dim = self.latent_dim - self.expert_dim
        self.expanded = self.ml_dim > 0
        self.sparsity, self.output_sparsity = sparsity, output_sparsity
        self.t_max, self.step_size = t_max, step_size
        self.time_dim = int(t_max / step_size + 1)
        self.roche_config, self.dose_max, self.p_remove, self.output_sigma = roche_config, dose_max, p_remove, output_sigma
        self.exact_rng, self.rtol, self.atol, self._lib = bool(exact_rng), float(rtol), float(atol), lib
        # same draws, same order as dataloader.py:51-59
        n_in = self.latent_dim + self.action_dim
        self.output_coef = np.random.randn(obs_dim, n_in) * np.random.binomial(1, 1 - output_sparsity, (obs_dim, n_in))
        self.ml_coef = (np.random.randn(self.latent_dim, self.ml_dim)
                        * np.random.binomial(1, 1 - sparsity, (self.latent_dim, self.ml_dim)) / self.latent_dim)
        self.val_size, self.test_size = int(val_size), int(test_size)
        self.train_size = int(n_sample - val_size - test_size)
        self.measurements = self.actions = self.latents = self.masks = None
        self.dose_time = self.dose_amount = None
        self.data_train = self.data_val = self.data_test = None

    # ---- random inputs (dataloader.py:200-220) -------------------------------------------------------------------
    def get_initial_conditions(self):
        return np.random.exponential(scale=0.01, size=(self.n_sample, self.latent_dim))

    def get_action(self):
        days = np.stack([np.random.choice(self.t_max, size=1, replace=False) for _ in range(self.n_sample)], axis=0)
        return np.sort(days), np.random.rand(self.n_sample) * self.dose_max

    def _make_tensor(self, x):
        if isinstance(x, np.ndarray):
            return torch.tensor(x, dtype=self.dtype, device=self.device)
        return x.to(dtype=self.dtype, device=self.device)

    # ---- the solve: every patient in one launch ----------------------------------------------------------------------
    def true_parameters(self) -> torch.Tensor:
        """Packed parameter set of the generating ODE (layout of include/hode.h): the 13 ``roche_config`` scalars, then
        ``ml_coef.T`` as ``ml_net[0].weight`` and a zero bias (``dataloader.py:145`` has no bias term)."""
        parts = [np.asarray(list(self.roche_config), dtype=np.float64)]
        if self.expanded:
            parts += [self.ml_coef.T.reshape(-1), np.zeros(self.ml_dim)]
        return torch.tensor(np.concatenate(parts), dtype=torch.float32, device=self.device)[None].contiguous()

    def solve_latents(self, init, dose_time, dose_amount) -> torch.Tensor:
        """``init [N, D]``, ``dose_time [N, n_dose]``, ``dose_amount [N]`` (tensors on the cohort device) ->
        latents ``[time_dim, N, D]`` at ``t = 0, step, .., t_max`` (the grid of ``dataloader.py:155-160``)."""
        lib = self._lib
        if lib is None:
            if torch.device(self.device).type != "cuda":
                raise RuntimeError("DataGeneratorRoche integrates on CUDA (sm_100a) only; device is {}. There is no CPU "
                                   "fallback.".format(self.device))
            lib = L.get_lib()
        hc, hp = float(self.roche_config[0]), float(self.roche_config[1])
        n_dose = dose_time.shape[1]
        cfg = ops.make_cfg(L.FIELD_ROCHE, self.latent_dim, L.DOPRI5, controller=L.CTRL_TRAJ, n_dose=n_dose,
                           rtol=self.rtol, atol=self.atol, hill2=(hc == 2.0 and hp == 2.0))
        pb = ops.Problem(cfg, 1, init.shape[0], dose_amount.to(torch.float32).contiguous(),
                         dose_time.to(torch.float32).contiguous(), self.true_parameters(), None)
        t_eval = torch.arange(self.time_dim, dtype=torch.float64, device=init.device) * float(self.step_size)
        h, stats, _ = ops.dopri5_fwd(lib, pb, init.to(torch.float32).contiguous(), t_eval, 0)
        bad = torch.nonzero(stats[:, 3] != L.SOLVE_OK)
        if bad.numel():  # the reference's `while ode.successful()` would silently truncate; fail loudly instead
            raise RuntimeError("data generation: dopri5 failed for patient {} (status {})".format(
                int(bad[0]), int(stats[int(bad[0]), 3])))
        return h

    def generate_data(self):
        dev, T, N = self.device, self.time_dim, self.n_sample
        if self.exact_rng:
            init = self._make_tensor(self.get_initial_conditions())
            dose_time, dose_amount = self.get_action()
            noise = np.random.randn(N, self.obs_dim, T)  # == N successive randn(obs, T) calls (dataloader.py:171)
        else:
            g = torch.Generator(device=dev)
            g.manual_seed(int(np.random.randint(0, 2 ** 31 - 1)))
            init = torch.empty(N, self.latent_dim, device=dev, dtype=self.dtype).exponential_(100.0, generator=g)
            dose_time = torch.randint(0, int(self.t_max), (N, 1), device=dev, generator=g)
            dose_amount = torch.rand(N, device=dev, generator=g, dtype=torch.float64) * self.dose_max
        self.dose_time, self.dose_amount = dose_time, dose_amount
        dt_dev = torch.as_tensor(dose_time, device=dev)
        da_dev = torch.as_tensor(dose_amount, device=dev)
        self.latents = self.solve_latents(init, dt_dev, da_dev).to(self.dtype)
        # actions[t, i] = amount_i * [some dose of patient i is given at time t]   (dataloader.py:100-101, 179-183)
        times = torch.arange(T, device=dev, dtype=torch.float64) * float(self.step_size)
        hit = (dt_dev.to(torch.float64)[None, :, :] == times[:, None, None]).any(dim=-1)
        self.actions = (hit * da_dev.to(torch.float64)[None, :]).to(self.dtype)[..., None]
        # read-out + noise (dataloader.py:166-171)
        coef = torch.tensor(self.output_coef, device=dev)  # float64 [obs, D + 1]
        if self.exact_rng:
            out = self.latents.to(torch.float64) @ coef[:, :-1].T + coef[:, -1]
            out = out + torch.tensor(noise, device=dev).permute(2, 0, 1) * self.output_sigma
            measurements = out.to(self.dtype)
            # torch.rand_like fills in STORAGE order, and the reference's measurement tensor is a transposed view of
            # [N][obs][T] storage (dataloader.py:245): draw in that order, view time-major
            u = torch.rand(N, self.obs_dim, T, dtype=self.dtype).permute(2, 0, 1)
            selected = (u > self.p_remove).to(dev).contiguous() * 1.0
        else:
            c32 = coef.to(self.dtype)
            measurements = self.latents @ c32[:, :-1].T + c32[:, -1]
            measurements += torch.randn(T, N, self.obs_dim, device=dev, dtype=self.dtype, generator=g) * self.output_sigma
            selected = (torch.rand(T, N, self.obs_dim, device=dev, dtype=self.dtype, generator=g) > self.p_remove) * 1.0
        self.measurements = (measurements - torch.mean(measurements, dim=(0, 1))) / torch.std(measurements, dim=(0, 1))
        self.masks = torch.ones(T, N, 1, device=dev, dtype=self.dtype) * selected
        assert self.measurements.shape == (T, N, self.obs_dim)
        assert self.actions.shape == (T, N, self.action_dim)
        assert self.latents.shape == (T, N, self.latent_dim)

    # ---- splits and batches (dataloader.py:66-92, 268-342) ---------------------------------------------------------------
    def _fold(self, fold):
        assert fold in ("train", "val", "test")
        return {"train": self.data_train, "val": self.data_val, "test": self.data_test}[fold]

    def set_device(self, device):
        self.device = device
        for k in _KEYS:
            setattr(self, k, getattr(self, k).to(device))
        for d in (self.data_train, self.data_val, self.data_test):
            for k in _KEYS:
                d[k] = d[k].to(device)

    def split_sample(self):
        a, b = self.train_size, self.train_size + self.val_size
        cut = lambda lo, hi: {k: getattr(self, k)[:, lo:hi, :] for k in _KEYS}  # noqa: E731
        self.data_train, self.data_val, self.data_test = cut(0, a), cut(a, b), cut(b, None)

    def set_train_size(self, n_sample):
        self.train_size = n_sample - self.val_size - self.test_size
        self.n_sample = n_sample
        print("train_size", self.train_size)
        print("n_sample", self.n_sample)
        for k in _KEYS:
            self.data_train[k] = self.data_train[k][:, : self.train_size, :]

    def set_val_size(self, n_val):
        self.val_size = n_val
        for k in _KEYS:
            self.data_val[k] = self.data_val[k][:, :n_val, :]

    def get_mini_batch(self, fold, batch_size):
        data = self._fold(fold)
        n = data["measurements"].shape[1]
        idx = self._make_tensor(np.random.choice(n, batch_size, replace=False)).to(torch.int64)
        return {k: data[k][:, idx, :] for k in _KEYS}

    def get_split(self, fold, batch_size, chunk=0):
        data = self._fold(fold)
        lo, hi = chunk * batch_size, (chunk + 1) * batch_size
        return {k: data[k][:, lo:hi, :] for k in _KEYS}
