"""One training iteration pinned to the REFERENCE's own code: ``tests/golden/training_iter_d6.npz`` holds the loss, the
encoder output and every parameter gradient of ``VariationalInference.loss`` + ``backward`` computed by the reference's
``EncoderLSTM`` / ``RocheExpertDecoder`` / ``VariationalInference`` classes (``oracle/make_golden_training.py``; only the
missing torchdiffeq is restated).  The CPU test checks the fixture against the oracle decoder + the test's encoder replica
(same state_dict keys as the reference's); the GPU test runs the drop-in decoder with the fused loss against the fixture.
Tolerances: CPU 1e-5 (same arithmetic); GPU 1e-6 on the loss, 2e-5 on x_hat and 1e-4 norm-wise on the gradients (the fixture's
first step carries gradient -- SURVEY.md App. D.5 -- which the kernels' constant-first-step gradient differs from by 3.5e-5)."""
import os

import numpy as np
import pytest
import torch

import hybrid_ode_neurips_2021_b200 as H
from oracle import fields as OF

from _util import check as gate, relerr
from test_gpu_training import Encoder, iteration

GOLD = os.path.join(os.path.dirname(__file__), "golden", "training_iter_d6.npz")
D, OBS = 6, 20


def load():
    g = np.load(GOLD)
    sd = {p: {k[len(p) + 6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(p + "__sd__")} for p in ("enc", "dec")}
    grads = {p: {k[len(p) + 8:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(p + "__grad__")} for p in ("enc", "dec")}
    data = {"measurements": torch.from_numpy(g["x"]), "actions": torch.from_numpy(g["a"]), "masks": torch.from_numpy(g["mask"])}
    return g, sd, grads, data


def check(enc, dec, loss, grads, g, tol, tag="cpu", grad_tol=None):
    grad_tol = 5 * tol if grad_tol is None else grad_tol
    gate("training fixture [{}] loss".format(tag), abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])), tol)
    for name, ref in grads["enc"].items():
        got = dict(enc.named_parameters())[name].grad
        assert got is not None
        gate("training fixture [{}] grad encoder.{}".format(tag, name), relerr(got, ref), grad_tol)
    for name in ("ode.ml_net.0.weight", "ode.ml_net.0.bias", "output_function.0.weight", "output_function.0.bias"):
        got = dict(dec.named_parameters())[name].grad
        gate("training fixture [{}] grad decoder.{}".format(tag, name), relerr(got, grads["dec"][name]), grad_tol)


def test_fixture_is_reproduced_by_the_oracle_iteration_on_cpu():
    g, sd, grads, data = load()
    enc = Encoder(OBS, D, 2 * OBS)
    assert set(enc.state_dict().keys()) == set(sd["enc"].keys())  # the replica has the reference encoder's parameters
    enc.load_state_dict(sd["enc"])
    dec = OF.OracleDecoder(OBS, D, method="dopri5")
    dec.load_state_dict(sd["dec"])
    mu, _ = enc(data["measurements"], data["actions"], data["masks"])
    assert torch.allclose(mu, torch.from_numpy(g["mu"]), rtol=1e-6, atol=1e-8)
    x_hat, _ = dec(mu, data["actions"])
    loss = torch.sum((data["measurements"] - x_hat) ** 2 * data["masks"]) / data["measurements"].shape[1]
    loss.backward()
    assert relerr(x_hat, torch.from_numpy(g["x_hat"])) < 1e-5
    check(enc, dec, loss, grads, g, 1e-5)


@pytest.mark.gpu
def test_drop_in_iteration_matches_the_reference_fixture():
    dev = "cuda:0"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g, sd, grads, data = load()
    enc = Encoder(OBS, D, 2 * OBS).to(dev)
    enc.load_state_dict(sd["enc"])
    dec = H.RocheExpertDecoder(OBS, D, 1, 14, 1, method="dopri5", device=dev)
    dec.load_state_dict(sd["dec"])  # the reference's checkpoint keys load unchanged
    batch = {k: v.to(dev) for k, v in data.items()}
    mu, _ = enc(batch["measurements"], batch["actions"], batch["masks"])
    h = dec.solve(mu, batch["actions"])
    loss = H.masked_sse(dec, h, batch["measurements"], batch["masks"])
    loss.backward()
    gate("training fixture [gpu] x_hat", relerr(dec.output_function(h), torch.from_numpy(g["x_hat"])), 2e-5)
    # loss 1e-6.  Gradients 1e-4: the fixture was produced with torchdiffeq's differentiable first
    # step (the open point) while the kernels implement the constant-first-step gradient; the measured gap
    # is 3.5e-5 on the encoder gradients / ml_net bias and 3.6e-6 on the ml_net weights (B200, round 2).
    check(enc, dec, loss, grads, g, 1e-6, tag="gpu", grad_tol=1e-4)
