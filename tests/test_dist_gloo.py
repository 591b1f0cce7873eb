"""world_size-2 gloo test (CPU) of the data-parallel plumbing: contiguous sharding and the single packed
gradient all-reduce give the same parameter gradients and loss as one process on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hybrid_ode_neurips_2021_b200 import dist as hd
from oracle import fields as OF

from _util import make_cohort


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            blocks = [hd.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in blocks) == (n + w - 1) // w


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, flat=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    D, obs, B = 6, 20, 12
    torch.manual_seed(0)
    dec = OF.OracleDecoder(obs, D, method="rk4", options={"step_size": 0.25})
    y0, a, x, mask = make_cohort(B, D, obs=obs, seed=1)
    lo, hi = hd.shard_range(B, rank, world)
    params = [p for n, p in dec.named_parameters() if "ml_net" in n or "output_function" in n]
    fg = hd.FlatGrads(params, extra=1) if flat else None  # gradients as views of one flat buffer (bench.py's path)
    for _ in range(2 if flat else 1):  # the flat buffer is zeroed, not re-created, between steps
        if flat:
            fg.zero_()
        xh, _ = dec(y0[lo:hi], a[:, lo:hi])
        loss = torch.sum((x[:, lo:hi] - xh) ** 2 * mask[:, lo:hi]) / B  # global batch in the normalisation
        loss.backward()
    if flat:
        assert all(p.grad.data_ptr() >= fg.flat.data_ptr() and p.grad._base is fg.flat for p in params)  # still views
        fg.extra.copy_(loss.detach().reshape(1))
        total = fg.allreduce()
    else:
        total = hd.allreduce_grads(params, extra=loss.detach().reshape(1))
    if rank == 0:
        q.put((total.item(), [p.grad.clone() for p in params]))
    dist.destroy_process_group()


@pytest.mark.parametrize("flat", [False, True])
def test_two_rank_allreduce_equals_single_process(flat):
    D, obs, B = 6, 20, 12
    torch.manual_seed(0)
    dec = OF.OracleDecoder(obs, D, method="rk4", options={"step_size": 0.25})
    y0, a, x, mask = make_cohort(B, D, obs=obs, seed=1)
    xh, _ = dec(y0, a)
    loss = torch.sum((x - xh) ** 2 * mask) / B
    loss.backward()
    ref = [p.grad.clone() for n, p in dec.named_parameters() if "ml_net" in n or "output_function" in n]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, flat)) for r in range(2)]
    for p in procs:
        p.start()
    total, grads = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(total - loss.item()) <= 1e-5 * abs(loss.item())
    for g, r in zip(grads, ref):
        assert torch.allclose(g, r, rtol=1e-4, atol=1e-5 * float(r.abs().max()))
