"""Caller-side drop-ins around the hot path: the LSTM encoder and the variational
wrapper of ``/root/reference/model.py`` (``GaussianReparam`` 18-31, priors 34-45, ``EncoderLSTM`` 383-440,
``VariationalInference`` 1124-1214), with the same constructors, attributes, ``state_dict`` keys, checkpoint layout and
random-number consumption, so the reference's ``training_utils.variational_training_loop`` and ``experiments/run_simulation*.py``
run on them unchanged.

What changes underneath:

* ``EncoderLSTM.forward`` feeds the reversed, masked sequence to ``nn.LSTM`` in ONE call instead of ``T`` single-step calls
  from a Python loop (same recurrence, same result);
* ``VariationalInference.loss`` uses the fused read-out + masked SSE (``loss.masked_sse``: ``x_hat`` is never materialised)
  when the decoder is the drop-in ``RocheExpertDecoder`` on a CUDA device, and the reference's expression otherwise;
* ``mc_kl`` keeps the reference's sample-by-sample loop by default (identical random stream); ``mc_vectorised=True`` draws
  all ``mc_size`` samples at once (same estimator, different stream) and evaluates both log-densities in one pass.

These are plain PyTorch modules (no kernels of their own); the CUDA work stays in the decoder.
"""
from __future__ import annotations

import os

import torch
import torch.distributions as dist
from torch import nn

from .loss import masked_sse
from .model import DTYPE, _default_device

__all__ = ["GaussianReparam", "StandardNormalPrior", "ExponentialPrior", "EncoderLSTM", "VariationalInference"]


class GaussianReparam:
    """Diagonal Gaussian posterior: sampling by re-parameterisation and its log-density (model.py:18-31)."""

    @staticmethod
    def reparameterize(mu, log_var):
        std = torch.exp(0.5 * log_var)
        return torch.randn_like(std) * std + mu

    @staticmethod
    def log_density(mu, log_var, z):
        return dist.normal.Normal(mu, torch.exp(0.5 * log_var)).log_prob(z).sum(dim=-1)


class StandardNormalPrior:
    @staticmethod
    def log_density(z):
        zero, one = torch.tensor([0.0]).to(z), torch.tensor([1.0]).to(z)
        return dist.normal.Normal(zero, one).log_prob(z).sum(dim=-1)


class ExponentialPrior:
    """Exponential(rate = 100): the prior on the initial state, mean 0.01 like the generator's ``y0`` (model.py:41-45)."""

    @staticmethod
    def log_density(z):
        return dist.exponential.Exponential(rate=torch.tensor([100.0]).to(z)).log_prob(z).sum(dim=-1)


class EncoderLSTM(nn.Module, GaussianReparam):
    def __init__(self, input_dim, hidden_dim, output_dim, normalize=True, device=None):
        super().__init__()
        self.device = _default_device() if device is None else device
        self.hidden_dim = hidden_dim
        self.normalize = normalize
        self.model_name = "LSTMEncoder"
        self.lstm = nn.LSTM(input_dim, hidden_dim).to(self.device)
        self.lin = nn.Linear(hidden_dim, output_dim).to(self.device)       # posterior mean
        self.log_var = nn.Linear(hidden_dim, output_dim).to(self.device)   # posterior log-variance

    def forward(self, x, a, mask):
        """``x [T, B, obs]``, ``a [T, B, 1]``, ``mask [T, B, obs]`` -> ``mu, log_var [B, output_dim]``.  The reference runs the
        LSTM over ``t = T-1 .. 0`` one step at a time (model.py:424-426); one call on the flipped sequence is the same
        recurrence, and the state after the last step (``t = 0``) is what the two heads read."""
        if x.dim() == 4:  # model.py:412 `x.squeeze()`
            x = x.squeeze()
        seq = torch.cat([x, a], dim=-1) * torch.cat([mask, torch.ones_like(a)], dim=-1)
        out, _ = self.lstm(torch.flip(seq, dims=[0]))
        last = out[-1]
        mu, log_var = self.lin(last), self.log_var(last)
        if self.normalize:
            mu = torch.exp(mu) / 10
            log_var = log_var - 5.0
        return mu, log_var


class VariationalInference:
    epsilon = torch.finfo(DTYPE).eps

    def __init__(self, encoder, decoder, elbo=True, prior_log_pdf=None, mc_size=100, mc_vectorised=False):
        self.encoder, self.decoder = encoder, decoder
        self.prior_log_pdf, self.mc_size, self.elbo = prior_log_pdf, mc_size, elbo
        self.mc_vectorised = bool(mc_vectorised)
        self.model_name = "VI_{}_{}.pkl".format(encoder.model_name, decoder.model_name)

    def save(self, path, itr, best_loss):
        path = path + self.model_name
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save({"itr": itr, "encoder_state_dict": self.encoder.state_dict(),
                    "decoder_state_dict": self.decoder.state_dict(), "best_loss": best_loss}, path)

    def parameters(self):
        return list(self.encoder.parameters()) + list(self.decoder.parameters())

    def _fused(self, z):
        return z.is_cuda and hasattr(self.decoder, "solve") and hasattr(self.decoder, "output_function")

    def loss(self, data):
        x, a, mask = data["measurements"], data["actions"], data["masks"]
        self.x, self.a, self.mask = x, a, mask
        mu, log_var = self.encoder(x, a, mask)
        self.mu, self.log_var = mu, log_var
        z = self.encoder.reparameterize(mu, log_var) if self.elbo else mu
        self.z = z
        if self._fused(z):
            # the latent solve, then read-out + masked SSE in one streaming kernel; x_hat is not materialised
            self.h_hat = self.decoder.solve(z, a)
            self.x_hat = None
            lik = masked_sse(self.decoder, self.h_hat, x, mask)
        else:
            self.x_hat, self.h_hat = self.decoder(z, a)
            lik = torch.sum((x - self.x_hat) ** 2 * mask) / x.shape[1]
        if not self.elbo:
            return lik
        if self.prior_log_pdf is None:  # closed-form KL to a standard normal
            kld = torch.mean(-0.5 * torch.sum(1 + log_var - mu ** 2 - log_var.exp(), dim=1), dim=0)
        else:
            kld = torch.mean(self.mc_kl(mu, log_var, self.mc_size), dim=0)
        return lik + kld

    def mc_kl(self, mu, log_var, sample_size):
        """Monte-Carlo ``KL(q || p) = E_q[log q(z) - log p(z)]`` per patient, non-positive samples clamped to ``epsilon``
        (model.py:1198-1214)."""
        if self.mc_vectorised:
            std = torch.exp(0.5 * log_var)
            z = torch.randn((sample_size,) + tuple(std.shape), dtype=std.dtype, device=std.device) * std + mu
            z = torch.where(z <= 0.0, torch.full_like(z, self.epsilon), z)
            return (self.encoder.log_density(mu, log_var, z) - self.prior_log_pdf(z)).mean(dim=0)
        terms = []
        for _ in range(sample_size):
            z = self.encoder.reparameterize(mu, log_var)
            z[z <= 0.0] = self.epsilon
            terms.append(self.encoder.log_density(mu, log_var, z) - self.prior_log_pdf(z))
        return torch.stack(terms, dim=-1).mean(dim=-1)
