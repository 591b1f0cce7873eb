"""Data-parallel plumbing for the hot path: one process per GPU, trajectories (or ensemble members) sharded in
contiguous blocks, and ONE all-reduce per backward of the packed parameter-gradient vector.

The reference has no distributed code at all; the only quantity that couples trajectories across ranks is the sum over
the batch in the parameter gradients (``dL/dW_ml``, ``dL/db_ml``, ``dL/dW_out``, ``dL/db_out`` and the 13 expert scalars:
154 / 396 / 1 144 floats for D = 6 / 8 / 12) and the scalar loss, whose ``1/B`` uses the GLOBAL batch
(``model.py:1179``).  Forward-only sweeps need no communication.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``ceil(n / world)`` items owned by ``rank`` (last ranks may be short/empty)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def pack_grads(params: Iterable[torch.nn.Parameter]) -> Tuple[torch.Tensor, List[torch.nn.Parameter]]:
    ps = [p for p in params if p.requires_grad]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in ps])
    return flat, ps


def unpack_grads(flat: torch.Tensor, ps: List[torch.nn.Parameter]) -> None:
    off = 0
    for p in ps:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


def allreduce_grads(params: Iterable[torch.nn.Parameter], extra: torch.Tensor = None, group=None):
    """Sum the gradients of ``params`` (and an optional small tensor such as the local loss) over all ranks with a
    single collective.  Returns the reduced ``extra``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return extra
    flat, ps = pack_grads(params)
    n_extra = 0
    if extra is not None:
        n_extra = extra.numel()
        flat = torch.cat([flat, extra.reshape(-1).float()])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    unpack_grads(flat[: flat.numel() - n_extra], ps)
    return flat[flat.numel() - n_extra:].view_as(extra) if extra is not None else None


class FlatGrads:
    """Gradients of ``params`` kept as views of ONE flat float32 buffer: backward accumulates straight into the buffer, the
    all-reduce runs on it in place, and nothing is packed or unpacked (the ``torch.cat`` / slice-copy kernels of
    :func:`allreduce_grads` are ~10 tiny launches per step, which only matters once a step is ~1 ms: strong scaling).

    ``extra`` float slots at the end of the buffer carry small per-step scalars (the local loss) through the same collective.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], extra: int = 1):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGrads needs at least one trainable parameter")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n_extra = int(extra)
        self.flat = torch.zeros(n + self.n_extra, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("FlatGrads holds float32 gradients")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.n_params = n

    def zero_(self):
        self.flat.zero_()

    @property
    def grads(self) -> torch.Tensor:
        return self.flat[: self.n_params]

    @property
    def extra(self) -> torch.Tensor:
        return self.flat[self.n_params:]

    def allreduce(self, group=None) -> torch.Tensor:
        """Sum the buffer over all ranks in place (one collective); returns the reduced ``extra`` slots."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.extra
