"""Caller-side drop-ins (``vi.EncoderLSTM``, ``vi.VariationalInference``) against ``tests/golden/vi_elbo_d6.npz``, produced by
the reference's own classes with the stochastic ELBO terms under a fixed CPU seed (``oracle/make_golden_vi.py``).  The
drop-ins consume the random stream in the reference's order, so on the CPU -- with the oracle decoder standing in for the
CUDA decoder -- the sample ``z``, the Monte-Carlo KL, the loss and every gradient must be reproduced (1e-5)."""
import os

import numpy as np
import pytest
import torch

import hybrid_ode_neurips_2021_b200 as H
from hybrid_ode_neurips_2021_b200 import vi as V
from oracle import fields as OF

from _util import relerr

GOLD = os.path.join(os.path.dirname(__file__), "golden")
D, OBS = 6, 20


def load(name):
    g = np.load(os.path.join(GOLD, name))
    sd = {p: {k[len(p) + 6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(p + "__sd__")} for p in ("enc", "dec")}
    grads = {p: {k[len(p) + 8:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(p + "__grad__")} for p in ("enc", "dec")}
    data = {"measurements": torch.from_numpy(g["x"]), "actions": torch.from_numpy(g["a"]), "masks": torch.from_numpy(g["mask"])}
    return g, sd, grads, data


def build(sd, device="cpu"):
    enc = V.EncoderLSTM(OBS + 1, 2 * OBS, D, device=torch.device(device))
    assert list(enc.state_dict().keys()) == list(sd["enc"].keys())  # checkpoint-compatible with the reference encoder
    enc.load_state_dict(sd["enc"])
    dec = OF.OracleDecoder(OBS, D, method="dopri5")
    dec.model_name = "HybridDecoder"
    dec.load_state_dict(sd["dec"])
    return enc, dec


def test_stochastic_elbo_reproduces_the_reference_on_cpu():
    g, sd, grads, data = load("vi_elbo_d6.npz")
    enc, dec = build(sd)
    vi = V.VariationalInference(enc, dec, prior_log_pdf=V.ExponentialPrior.log_density, elbo=True, mc_size=int(g["mc_size"]))
    assert vi.model_name == str(g["model_name"])
    torch.manual_seed(int(g["seed"]))
    loss = vi.loss(data)
    loss.backward()
    assert torch.allclose(vi.mu, torch.from_numpy(g["mu"]), rtol=1e-6, atol=1e-8)
    assert torch.allclose(vi.log_var, torch.from_numpy(g["log_var"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(vi.z, torch.from_numpy(g["z"]), rtol=1e-6, atol=1e-8)  # same random stream
    assert abs(loss.item() - float(g["loss_mc"])) <= 1e-5 * abs(float(g["loss_mc"]))
    for name, ref in grads["enc"].items():
        assert relerr(dict(enc.named_parameters())[name].grad, ref) < 5e-5, name
    for name in ("ode.ml_net.0.weight", "output_function.0.weight"):
        assert relerr(dict(dec.named_parameters())[name].grad, grads["dec"][name]) < 5e-5, name
    # closed-form KL variant
    vi2 = V.VariationalInference(enc, dec, prior_log_pdf=None, elbo=True)
    torch.manual_seed(int(g["seed"]))
    assert abs(vi2.loss(data).item() - float(g["loss_closed_form"])) <= 1e-5 * abs(float(g["loss_closed_form"]))


def test_vectorised_mc_kl_is_the_same_estimator():
    g, sd, _, data = load("vi_elbo_d6.npz")
    enc, dec = build(sd)
    mu, log_var = (torch.from_numpy(g[k]) for k in ("mu", "log_var"))
    loop = V.VariationalInference(enc, dec, prior_log_pdf=V.ExponentialPrior.log_density, mc_size=4000)
    vec = V.VariationalInference(enc, dec, prior_log_pdf=V.ExponentialPrior.log_density, mc_size=4000, mc_vectorised=True)
    torch.manual_seed(0)
    k_loop = loop.mc_kl(mu, log_var, 4000)
    torch.manual_seed(1)
    k_vec = vec.mc_kl(mu, log_var, 4000)
    assert k_loop.shape == k_vec.shape == (mu.shape[0],)
    assert relerr(k_vec, k_loop) < 0.05  # two independent 4000-sample estimates of the same expectation
    # gradients flow to the posterior parameters through the re-parameterised samples, clamped samples excluded
    m = mu.clone().requires_grad_(True)
    vec.mc_kl(m, log_var, 64).sum().backward()
    assert m.grad is not None and bool(torch.isfinite(m.grad).all())


def test_fused_branch_and_checkpoint_layout(tmp_path, monkeypatch):
    """The CUDA branch of ``loss`` (decoder.solve + fused masked SSE) executed on the CPU with the kernel call replaced by its
    definition: same value as the generic branch; ``save`` writes the reference's checkpoint layout."""
    g, sd, _, data = load("training_iter_d6.npz")
    enc, dec = build(sd)

    class WithSolve(OF.OracleDecoder):
        def solve(self, init, a):
            return self(init, a)[1]

    dec.__class__ = WithSolve
    vi = V.VariationalInference(enc, dec, elbo=False)
    generic = vi.loss(data)
    assert abs(generic.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    monkeypatch.setattr(V.VariationalInference, "_fused", lambda self, z: True)
    monkeypatch.setattr(V, "masked_sse", lambda d, h, x, m: torch.sum((x - d.output_function(h)) ** 2 * m) / x.shape[1])
    fused = vi.loss(data)
    assert vi.x_hat is None and torch.allclose(fused, generic, rtol=1e-6)
    vi.save(str(tmp_path) + "/", 3, 1.5)
    ck = torch.load(os.path.join(str(tmp_path), vi.model_name))
    assert set(ck) == {"itr", "encoder_state_dict", "decoder_state_dict", "best_loss"} and ck["itr"] == 3
    assert len(vi.parameters()) == len(list(enc.parameters())) + len(list(dec.parameters()))


@pytest.mark.gpu
def test_variational_inference_drop_in_on_gpu_matches_the_reference_fixture():
    dev = "cuda:0"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g, sd, grads, data = load("training_iter_d6.npz")
    enc = V.EncoderLSTM(OBS + 1, 2 * OBS, D, device=torch.device(dev))
    enc.load_state_dict(sd["enc"])
    dec = H.RocheExpertDecoder(OBS, D, 1, 14, 1, method="dopri5", device=dev)
    dec.load_state_dict(sd["dec"])
    vi = V.VariationalInference(enc, dec, elbo=False)
    loss = vi.loss({k: v.to(dev) for k, v in data.items()})
    loss.backward()
    assert vi.x_hat is None  # fused path taken
    assert abs(loss.item() - float(g["loss"])) <= 1e-3 * abs(float(g["loss"]))
    for name, ref in grads["enc"].items():
        assert relerr(dict(enc.named_parameters())[name].grad, ref) < 5e-3, name


def test_reference_training_loop_runs_unmodified_on_the_drop_ins(tmp_path):
    """The reference's own ``training_utils.variational_training_loop`` (imported unmodified; build container only) drives
    the drop-in generator (through the host emulation), encoder and VariationalInference for a few iterations with the oracle
    decoder standing in for the CUDA decoder: API compatibility of the callers' side (mini-batches, loss, save / reload of
    the checkpoint layout)."""
    import subprocess

    from hybrid_ode_neurips_2021_b200 import _lib as L
    from hybrid_ode_neurips_2021_b200.datagen import DataGeneratorRoche
    from hybrid_ode_neurips_2021_b200.model import RochConfig
    from oracle import refload

    if not refload.available():
        pytest.skip("/root/reference not present (GPU box)")
    TU = refload.load("training_utils")
    hs_dir = os.path.join(os.path.dirname(__file__), "hostsim")
    subprocess.run(["make", "-C", hs_dir], check=True, capture_output=True)
    lib = L.HodeLib(os.path.join(hs_dir, "libhode_hostsim.so"), required=["hode_abi_version", "hode_last_error", "hode_dopri5_fwd"])
    np.random.seed(666)
    torch.manual_seed(666)
    dg = DataGeneratorRoche(40, OBS, 14, 1, RochConfig(kel=1), 0.1, 1, D, 0.5, p_remove=0.5, output_sparsity=0.5,
                            device=torch.device("cpu"), val_size=8, test_size=8, lib=lib)
    dg.generate_data()
    dg.split_sample()
    enc = V.EncoderLSTM(OBS + 1, 2 * OBS, D, device=torch.device("cpu"))
    dec = OF.OracleDecoder(OBS, D, method="rk4", options={"step_size": 0.125})
    dec.model_name = "HybridDecoder"
    vi = V.VariationalInference(enc, dec, prior_log_pdf=V.ExponentialPrior.log_density, elbo=True, mc_size=5)
    params = list(enc.parameters()) + list(dec.output_function.parameters()) + list(dec.ode.ml_net.parameters())
    opt = torch.optim.Adam(params, lr=0.01)
    model, best, seconds = TU.variational_training_loop(niters=4, data_generator=dg, model=vi, batch_size=8, optimizer=opt,
                                                        test_freq=2, path=str(tmp_path) + "/", shuffle=True)
    assert model is vi and np.isfinite(best) and best < 1e9
    assert os.path.exists(os.path.join(str(tmp_path), vi.model_name))
