"""Drop-ins for ``training_utils.evaluate`` / ``evaluate_horizon`` / ``evaluate_ensemble`` / ``evaluate_ensemble_horizon``
(``hybrid_ode_neurips_2021_b200.evaluation``) against the reference's own functions (``baseline/_ref/training_utils.py``,
imported unmodified): same models, same cohort, same seeds.

* ``evaluate`` / ``evaluate_ensemble``: both sides run on the GPU with identical random streams (the drop-ins draw in the
  reference's order), so all six returned statistics must agree to rounding -- the reference through ~10^4 ``.item()`` /
  ``properscoring`` calls, the drop-in through one solve launch and one CRPS launch per chunk.
* ``evaluate_horizon`` / ``evaluate_ensemble_horizon``: the reference calls ``.numpy()`` on its results, i.e. it only runs on
  host tensors; it is run on the CPU (oracle solver) with the posterior variance switched off, which makes the Monte-Carlo
  part independent of the random stream, and compared with the drop-in on the GPU.
"""
import contextlib
import io
import sys

import numpy as np
import pytest
import torch

import hybrid_ode_neurips_2021_b200 as H
from oracle import fields as OF
from oracle import refload

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
D, OBS = 6, 20


def _models(M, device, decoder_cls, expert_only=False, neural=False, seed=0, **dec_kw):
    torch.manual_seed(seed)
    d = 4 if expert_only else D
    enc = M.EncoderLSTM(OBS + 1, 2 * OBS, d, device=torch.device(device), normalize=not neural)
    dec = decoder_cls(OBS, d, 1, 14, 1, roche=not neural, device=torch.device(device), **dec_kw)
    return M.VariationalInference(enc, dec, prior_log_pdf=None if neural else M.ExponentialPrior.log_density, elbo=True)


def _cohort(n=50, val=10, test=20):
    np.random.seed(666)
    torch.manual_seed(666)
    dg = H.DataGeneratorRoche(n, OBS, 14, 1, H.RochConfig(kel=1), 0.2, 10, D, 0.5, p_remove=0.5, output_sparsity=0.5,
                              device=torch.device(DEV), val_size=val, test_size=test)
    dg.generate_data()
    dg.split_sample()
    return dg


def _quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


@pytest.fixture(scope="module")
def ref_modules():
    if not refload.available():
        pytest.skip("reference tree not available (baseline/_ref missing)")
    sys.modules.pop("torchdiffeq", None)
    H.install_as_torchdiffeq(force=True)
    refload.install_shims()
    for name in ("model", "training_utils"):
        sys.modules.pop(name, None)
    return refload.load("model"), refload.load("training_utils")


def test_evaluate_and_evaluate_ensemble_equal_the_reference_functions(ref_modules):
    M, TU = ref_modules
    dg = _cohort()
    vi = _models(M, DEV, H.RocheExpertDecoder, method="dopri5")          # reference encoder / VI, drop-in decoder
    vi_e = _models(M, DEV, H.RocheExpertDecoder, expert_only=True, seed=1, method="dopri5")
    vi_n = _models(M, DEV, H.RocheExpertDecoder, neural=True, seed=2, method="dopri5")
    for ref_fn, our_fn, args in ((TU.evaluate, H.evaluate, (vi,)), (TU.evaluate_ensemble, H.evaluate_ensemble, (vi_e, vi_n))):
        kw = dict(mc_itr=6) if len(args) == 1 else dict(mc_itr=6, weight_expert=0.7, weight_ml=0.4)
        torch.manual_seed(11)
        ref, ref_out = _quiet(ref_fn, *args, dg, 10, 5, **kw)
        torch.manual_seed(11)
        got, got_out = _quiet(our_fn, *args, dg, 10, 5, **kw)
        assert len(ref) == len(got) == 6
        for name, r, g in zip(("rmse_z0", "rmse_z0_sd", "cprs_z0", "rmse_x", "rmse_x_sd", "cprs_x"), ref, got):
            assert abs(g - r) <= 2e-4 * abs(r) + 1e-7, (name, r, g)
        assert [ln.split(",")[0] for ln in got_out.splitlines()] == [ln.split(",")[0] for ln in ref_out.splitlines()]


def test_horizon_variants_against_the_reference_on_the_cpu(ref_modules):
    M, TU = ref_modules
    dg = _cohort(n=30, val=6, test=8)
    pair = []
    for kind, seed in (("expert", 3), ("neural", 4)):
        vi = _models(M, DEV, H.RocheExpertDecoder, expert_only=kind == "expert", neural=kind == "neural", seed=seed, method="dopri5")
        with torch.no_grad():  # posterior std -> exp(-40): every Monte-Carlo sample equals the posterior mean
            vi.encoder.log_var.weight.zero_()
            vi.encoder.log_var.bias.fill_(-80.0)
        pair.append(vi)
    got = H.evaluate_horizon(pair[0], dg, 4, 5, mc_itr=3)
    got_e = H.evaluate_ensemble_horizon(pair[0], pair[1], dg, 4, 5, mc_itr=3, weight_expert=0.6, weight_ml=0.5)
    # the reference functions on host tensors, with the oracle solver under the reference's decoder classes
    sys.modules.pop("torchdiffeq", None)
    refload.install_shims()
    for name in ("model", "training_utils"):
        sys.modules.pop(name, None)
    Mc, TUc = refload.load("model"), refload.load("training_utils")
    dg.set_device(torch.device("cpu"))
    cpu = []
    for vi, kind in zip(pair, ("expert", "neural")):
        c = _models(Mc, "cpu", Mc.RocheExpertDecoder, expert_only=kind == "expert", neural=kind == "neural", method="dopri5")
        c.encoder.load_state_dict({k: v.cpu() for k, v in vi.encoder.state_dict().items()})
        c.decoder.load_state_dict({k: v.cpu() for k, v in vi.decoder.state_dict().items()})
        cpu.append(c)
    ref = TUc.evaluate_horizon(cpu[0], dg, 4, 5, mc_itr=3)
    ref_e = TUc.evaluate_ensemble_horizon(cpu[0], cpu[1], dg, 4, 5, mc_itr=3, weight_expert=0.6, weight_ml=0.5)
    for g, r in ((got, ref), (got_e, ref_e)):
        assert set(g) == set(r) == {"rmse_x", "rmse_x_sd", "cprs_x", "cprs_x_sd"}
        for key in ("rmse_x", "cprs_x", "cprs_x_sd"):  # (rmse_x_sd is a bootstrap: different draws on the two sides)
            assert g[key].shape == r[key].shape
            assert np.allclose(g[key], r[key], rtol=2e-3, atol=1e-5, equal_nan=True), key
    sys.modules.pop("torchdiffeq", None)
