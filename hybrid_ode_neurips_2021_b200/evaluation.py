"""Batched Monte-Carlo evaluation of a decoder: what ``training_utils.evaluate`` / ``evaluate_horizon`` /
``evaluate_ensemble`` (``/root/reference/training_utils.py:100-201, 204-330, 380-470``) do with ``mc_itr = 50`` separate
decoder solves per test chunk and ``properscoring.crps_ensemble`` called from Python double / triple loops.

Here the ``mc_itr`` solves are ONE launch (every sample is its own controller group, i.e. exactly the ``mc_itr``
independent ``odeint`` calls of ``training_utils.py:144-151``), and the CRPS is computed on the device -- fused with the
read-out, so the ``[T, B, obs, mc]`` prediction tensor the reference stacks (``:165``) is never materialised.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import ops

__all__ = ["crps_ensemble", "mc_solve", "decode_crps", "evaluate_chunk", "evaluate", "evaluate_horizon", "evaluate_ensemble",
           "evaluate_ensemble_horizon", "bootstrap_RMSE"]


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


# ----------------------------------------------------------------------------------------------------------------------
# Drop-ins for training_utils.evaluate / evaluate_horizon / evaluate_ensemble / evaluate_ensemble_horizon: same arguments,
# same printed lines, same return values, same consumption of the random streams (one ``reparameterize`` per Monte-Carlo
# iteration, in the reference's order; the 500 ``torch.randint`` draws of ``bootstrap_RMSE``).  What changes underneath:
#   * the mc_itr decoder solves of a chunk are ONE launch when the decoder is the drop-in (``mc_solve``);
#   * CRPS is computed on the device (``crps_ensemble`` / fused ``decode_crps``) instead of ``properscoring`` being called
#     once per scalar from double / triple Python loops with a ``.item()`` and a ``.cpu().numpy()`` each;
#   * ``bootstrap_RMSE`` draws its 500 index vectors in one call (the same numbers) and reads the result back once.
# ----------------------------------------------------------------------------------------------------------------------
def bootstrap_RMSE(err_sq):
    """``training_utils.bootstrap_RMSE`` (``:568-577``): std of 500 bootstrap RMSEs, same CPU index stream."""
    if isinstance(err_sq, np.ndarray):
        err_sq = torch.tensor(err_sq)
    n = len(err_sq)
    idx = torch.randint(n, (500,) + tuple(err_sq.shape))  # == 500 consecutive torch.randint(n, err_sq.shape) calls
    res = torch.sqrt(torch.mean(err_sq[idx.to(err_sq.device)], dim=tuple(range(1, err_sq.dim() + 1))))
    return np.std(np.array(res.cpu().tolist()))


def _one_launch(decoder, statics):
    return statics is None and hasattr(decoder, "solve") and hasattr(decoder, "output_function")


def _sample_predictions(decoder, z_list, actions, statics, t0):
    """Predictions of the MC samples for ``t >= t0``: ``[T - t0, B, obs, mc]`` (what the reference stacks, ``:165``)."""
    mc, B = len(z_list), z_list[0].shape[0]
    if _one_launch(decoder, statics):
        h = mc_solve(decoder, torch.stack(z_list), actions)  # [T, mc * B, D], one launch
        xh = decoder.output_function(h[t0:])
        return xh.reshape(xh.shape[0], mc, B, xh.shape[-1]).permute(0, 2, 3, 1)
    outs = [(decoder(z, actions, statics) if statics is not None else decoder(z, actions))[0] for z in z_list]
    return torch.stack(outs, dim=-1)[t0:]


def _chunk(models, weights, data, t0, mc_itr, real, expert_dim, want_z):
    """One test chunk: point-estimate error terms and Monte-Carlo CRPS for one model or a weighted pair of models."""
    x, a, mask = data["measurements"][:t0], data["actions"][:t0], data["masks"][:t0]
    z0 = data["latents"][0]
    statics = data["statics"] if real else None
    enc_out, x_hat = [], None
    for mdl, w in zip(models, weights):
        a_in = torch.cat([a, data["statics"][:t0]], dim=-1) if real else a
        eo = mdl.encoder(x, a_in, mask)
        xh = (mdl.decoder(eo[0], data["actions"], statics) if real else mdl.decoder(eo[0], data["actions"]))[0]
        xh = xh if len(models) == 1 else xh * w
        x_hat = xh if x_hat is None else x_hat + xh
        enc_out.append(eo)
    x_hat = x_hat[t0:]
    x_test, mask_test = data["measurements"][t0:], data["masks"][t0:]
    z0_hat = enc_out[0][0]
    # Monte-Carlo samples: the reference alternates reparameterize / decode per iteration (and expert / ml per model);
    # decoding consumes no random numbers, so drawing all samples first leaves every stream unchanged
    z_lists = [[] for _ in models]
    for _ in range(mc_itr):
        for k, mdl in enumerate(models):
            z_lists[k].append(mdl.encoder.reparameterize(*enc_out[k]))
    single = len(models) == 1
    if single and _one_launch(models[0].decoder, statics) and len(models[0].decoder.output_function) == 1:
        h = mc_solve(models[0].decoder, torch.stack(z_lists[0]), data["actions"])
        crps_x = decode_crps(models[0].decoder, h[t0:], x_test, mc_itr)  # read-out fused with the CRPS: [T - t0, B, obs]
    else:
        pred = None
        for k, (mdl, w) in enumerate(zip(models, weights)):
            p = _sample_predictions(mdl.decoder, z_lists[k], data["actions"], statics, t0)
            p = p if single else p * w
            pred = p if pred is None else pred + p
        crps_x = crps_ensemble(x_test, pred)
    out = {"x_test": x_test, "mask_test": mask_test, "x_hat": x_hat, "crps_x": crps_x}
    if want_z:
        out["se_z0"] = torch.sum((z0[:, :expert_dim] - z0_hat[:, :expert_dim]) ** 2, dim=1)
        z_mat = torch.stack(z_lists[0], dim=-1)[:, :expert_dim]  # B, expert_dim, MC
        out["crps_z0"] = crps_ensemble(z0[:, :expert_dim], z_mat).mean(dim=1)
    return out


def _evaluate(models, weights, data_generator, batch_size, t0, mc_itr, real):
    with torch.no_grad():
        se_z0, se_x, c_z0, c_x = [], [], [], []
        for chunk in range(data_generator.test_size // batch_size):
            r = _chunk(models, weights, data_generator.get_split("test", batch_size, chunk), t0, mc_itr, real,
                       data_generator.expert_dim, True)
            se_z0.append(r["se_z0"])
            se_x.append(torch.sum((r["x_test"] - r["x_hat"]) ** 2 * r["mask_test"], dim=(0, 2)) / torch.sum(r["mask_test"], dim=(0, 2)))
            c_z0.append(r["crps_z0"].double().cpu().numpy())
            c_x.append(r["crps_x"].double().mean(dim=(0, 2)).cpu().numpy())
        se_z0 = torch.cat(se_z0)
        rmse_z0, rmse_z0_sd = torch.sqrt(torch.mean(se_z0)).item(), bootstrap_RMSE(se_z0)
        c_z0 = np.concatenate(c_z0)
        cprs_z0, cprs_z0_sd = np.mean(c_z0), np.std(c_z0) / np.sqrt(len(c_z0))
        se_x = torch.cat(se_x)
        if len(models) == 1:  # evaluate() drops patients without a single observed forecast entry; evaluate_ensemble does not
            se_x = se_x[~torch.isnan(se_x)]
        rmse_x, rmse_x_sd = torch.sqrt(torch.mean(se_x)).item(), bootstrap_RMSE(se_x)
        c_x = np.concatenate(c_x)
        cprs_x, cprs_x_sd = np.mean(c_x), np.std(c_x) / np.sqrt(len(c_x))
        print("rmse_z0,{:.4f},{:.4f}".format(rmse_z0, rmse_z0_sd))
        print("rmse_x,{:.4f},{:.4f}".format(rmse_x, rmse_x_sd))
        print("cprs_z0,{:.4f},{:.4f}".format(cprs_z0, cprs_z0_sd))
        print("cprs_x,{:.4f},{:.4f}".format(cprs_x, cprs_x_sd))
        return rmse_z0, rmse_z0_sd, cprs_z0, rmse_x, rmse_x_sd, cprs_x


def _evaluate_horizon(models, weights, data_generator, batch_size, t0, mc_itr, real, first_chunk_only):
    with torch.no_grad():
        se_x, c_x = [], []
        for chunk in range(data_generator.test_size // batch_size):
            r = _chunk(models, weights, data_generator.get_split("test", batch_size, chunk), t0, mc_itr, real,
                       data_generator.expert_dim, False)
            se_x.append(torch.sum((r["x_test"] - r["x_hat"]) ** 2 * r["mask_test"], dim=2) / torch.sum(r["mask_test"], dim=2))
            c_x.append(r["crps_x"].double().mean(dim=2).cpu().numpy())
            if first_chunk_only:
                break
        se_x = torch.cat(se_x, dim=1).cpu()  # T, B  (the reference calls .numpy() on it: host tensors)
        rmse_x = torch.sqrt(torch.nanmean(se_x, dim=1)).numpy()
        rmse_x_sd = np.array([bootstrap_RMSE(se_x[i]) for i in range(rmse_x.shape[0])])
        c_x = np.concatenate(c_x, axis=1)
        return {"rmse_x": rmse_x, "rmse_x_sd": rmse_x_sd, "cprs_x": np.mean(c_x, axis=1),
                "cprs_x_sd": np.std(c_x, axis=1) / np.sqrt(c_x.shape[1])}


def evaluate(model, data_generator, batch_size, t0, mc_itr=50, real=False):
    """``training_utils.evaluate`` (``:100-201``)."""
    return _evaluate([model], [1], data_generator, batch_size, t0, mc_itr, real)


def evaluate_horizon(model, data_generator, batch_size, t0, mc_itr=10, real=False):
    """``training_utils.evaluate_horizon`` (``:204-280``): per-time RMSE / CRPS of the forecast."""
    return _evaluate_horizon([model], [1], data_generator, batch_size, t0, mc_itr, real, False)


def evaluate_ensemble(model_expert, model_ml, data_generator, batch_size, t0, mc_itr=50, weight_expert=1, weight_ml=1):
    """``training_utils.evaluate_ensemble`` (``:383-487``): the weighted sum of an expert-ODE and a neural-ODE model (weights
    scalars or the per-time NNLS weights of ``run_simulation_ensemble.py:130-138``); latent metrics from the expert model."""
    return _evaluate([model_expert, model_ml], [weight_expert, weight_ml], data_generator, batch_size, t0, mc_itr, False)


def evaluate_ensemble_horizon(model_expert, model_ml, data_generator, batch_size, t0, mc_itr=10, weight_expert=1, weight_ml=1):
    """``training_utils.evaluate_ensemble_horizon`` (``:490-565``).  The reference aggregates and returns INSIDE its chunk loop
    (indentation of ``:545-565``), i.e. it reports the first test chunk only; reproduced as is."""
    return _evaluate_horizon([model_expert, model_ml], [weight_expert, weight_ml], data_generator, batch_size, t0, mc_itr,
                             False, True)
