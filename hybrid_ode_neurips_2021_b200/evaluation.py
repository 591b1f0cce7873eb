"""Batched Monte-Carlo evaluation of a decoder: what ``training_utils.evaluate`` / ``evaluate_horizon`` /
``evaluate_ensemble`` (``/root/reference/training_utils.py:100-201, 204-330, 380-470``) do with ``mc_itr = 50`` separate
decoder solves per test chunk and ``properscoring.crps_ensemble`` called from Python double / triple loops.

Here the ``mc_itr`` solves are ONE launch (every sample is its own controller group, i.e. exactly the ``mc_itr``
independent ``odeint`` calls of ``training_utils.py:144-151``), and the CRPS is computed on the device -- fused with the
read-out, so the ``[T, B, obs, mc]`` prediction tensor the reference stacks (``:165``) is never materialised.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops

__all__ = ["crps_ensemble", "mc_solve", "decode_crps", "evaluate_chunk"]


def crps_ensemble(observations: torch.Tensor, forecasts: torch.Tensor) -> torch.Tensor:
    """``properscoring.crps_ensemble(observations, forecasts)`` for equally weighted members on the last axis."""
    if not observations.is_cuda:
        raise RuntimeError("crps_ensemble runs on CUDA (sm_100a) only; there is no CPU fallback")
    return ops.crps_ensemble(L.get_lib(), observations.float(), forecasts.float())


def mc_solve(decoder, z_samples: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    """``z_samples [mc, B, D]`` (``mc`` draws of ``encoder.reparameterize``), ``actions [T, B, 1]`` ->
    ``h [T, mc * B, D]``, trajectory ``s * B + b`` = draw ``s`` of patient ``b``."""
    mc, B, D = z_samples.shape
    opts = dict(decoder.solver_options or {})
    opts["n_groups"] = mc * int(opts.get("n_groups", 1))
    saved = decoder.solver_options
    decoder.solver_options = opts
    try:
        return decoder.solve(z_samples.reshape(mc * B, D), actions.repeat(1, mc, 1))
    finally:
        decoder.solver_options = saved


def decode_crps(decoder, h: torch.Tensor, x: torch.Tensor, mc: int) -> torch.Tensor:
    """CRPS of the ``mc`` read-outs ``output_function(h[t, s*B + b])`` against ``x[t, b, o]`` -> ``[T, B, obs]``."""
    lin = decoder.output_function[0]
    return ops.decode_crps(L.get_lib(), h, lin.weight.detach(), lin.bias.detach(), x.float(), mc)


@torch.no_grad()
def evaluate_chunk(decoder, z0_hat, z_samples, actions, x, mask, t0: int):
    """One test chunk of ``training_utils.evaluate`` (``:108-178``) after the encoder: point-estimate squared error and
    MC CRPS of the forecast ``x[t0:]``.  Returns per-patient ``rmse_x`` terms (``sum((x - x_hat)^2 mask) / sum(mask)``,
    ``:137-139``) and per-patient ``crps_x`` (mean over time and observation, ``:176``)."""
    mc = z_samples.shape[0]
    x_hat, _ = decoder(z0_hat, actions)
    x_hat = x_hat[t0:]
    x_test, m_test = x[t0:], mask[t0:]
    se = torch.sum((x_test - x_hat) ** 2 * m_test, dim=(0, 2)) / torch.sum(m_test, dim=(0, 2))
    h = mc_solve(decoder, z_samples, actions)
    crps = decode_crps(decoder, h[t0:], x_test, mc)
    return {"se_x": se, "crps_x": crps.mean(dim=(0, 2)), "crps_x_full": crps}
