"""Fused read-out + masked squared error: ``output_function`` (``model.py:1097-1100, 1120``) followed by the
likelihood term of ``VariationalInference.loss`` (``model.py:1179``), ``sum((x - (W h + b))**2 * mask) / B``.

One streaming pass over ``x`` and ``mask`` produces the loss AND its gradients w.r.t. ``h``, ``W`` and ``b`` (the loss
is a scalar, so backward is a scaling of those) -- ``x_hat`` is never materialised.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops

__all__ = ["masked_sse", "decode_sse_loss"]


class _DecodeSSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, W, b, x, mask, n_norm):
        need = torch.is_grad_enabled() and (h.requires_grad or W.requires_grad or b.requires_grad)
        loss, gh, gw, gb = ops.decode_sse(L.get_lib(), h.detach(), W.detach(), b.detach(), x, mask, n_norm, True)
        ctx.save_for_backward(gh, gw, gb)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        gh, gw, gb = ctx.saved_tensors
        return g * gh, g * gw, g * gb, None, None, None


def decode_sse_loss(h, weight, bias, x, mask, n_norm=None):
    """``sum((x - (h @ weight.T + bias))**2 * mask) / n_norm`` with ``n_norm = x.shape[1]`` by default."""
    if not h.is_cuda:
        raise RuntimeError("decode_sse_loss runs on CUDA (sm_100a) only; there is no CPU fallback")
    if x.stride() != mask.stride():
        mask = mask.contiguous()
        x = x.contiguous()
    return _DecodeSSE.apply(h, weight, bias, x.float(), mask.float(), float(x.shape[1] if n_norm is None else n_norm))


def masked_sse(decoder, h, x, mask, n_norm=None):
    """Likelihood of ``VariationalInference.loss`` for a decoder whose ``output_function`` is ``Sequential(Linear)``."""
    lin = decoder.output_function[0]
    return decode_sse_loss(h, lin.weight, lin.bias, x, mask, n_norm)
