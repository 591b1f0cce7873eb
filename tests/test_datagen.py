"""Cohort generator: ``datagen.DataGeneratorRoche`` against fixtures produced by the REFERENCE's own
``dataloader.DataGeneratorRoche`` (``oracle/make_golden_datagen.py``: scipy lsoda per patient, seed 666).

With the same seeds every random quantity must be bit-identical (coefficients, doses, actions, masks); latents agree to
5e-5 absolute on states up to ~4 (float32 dopri5 at 1e-7/1e-8 vs float64 lsoda; measured 1.3e-5), and the normalised
measurements to 1e-4 (measured 2.7e-5).
The CPU run drives the kernel SOURCE through tests/hostsim (test-only emulation); the GPU run is the product path."""
import os
import subprocess

import numpy as np
import pytest
import torch

from hybrid_ode_neurips_2021_b200 import _lib as L
from hybrid_ode_neurips_2021_b200.datagen import DataGeneratorRoche
from hybrid_ode_neurips_2021_b200.model import RochConfig

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HS_DIR = os.path.join(os.path.dirname(__file__), "hostsim")
HS = os.path.join(HS_DIR, "libhode_hostsim.so")
CASES = ["datagen_d4", "datagen_d6", "datagen_d8"]


def build(name, device, lib=None, **kw):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    n, obs, D, val, test = (int(v) for v in g["cfg"])
    sigma, dose_max, sparsity, out_sparsity, p_remove = (float(v) for v in g["cfg_f"])
    np.random.seed(666)
    torch.manual_seed(666)
    dg = DataGeneratorRoche(n, obs, 14, 1, RochConfig(kel=1), sigma, dose_max, D, sparsity, p_remove=p_remove,
                            output_sparsity=out_sparsity, device=device, val_size=val, test_size=test, lib=lib, **kw)
    return g, dg


def check(g, dg):
    assert np.array_equal(dg.output_coef, g["output_coef"]) and np.array_equal(dg.ml_coef, g["ml_coef"])
    assert np.array_equal(dg.dose_time, g["dose_time"]) and np.array_equal(dg.dose_amount, g["dose_amount"])
    assert torch.equal(dg.actions.cpu(), torch.from_numpy(g["actions"]))  # bit-exact dose indexing
    assert torch.equal(dg.masks.cpu(), torch.from_numpy(g["masks"]))      # bit-exact masks
    lat, ref = dg.latents.cpu(), torch.from_numpy(g["latents"])
    assert lat.shape == ref.shape and (lat - ref).abs().max() < 5e-5, float((lat - ref).abs().max())
    assert (dg.measurements.cpu() - torch.from_numpy(g["measurements"])).abs().max() < 1e-4
    dg.split_sample()
    assert torch.allclose(dg.data_train["latents"].cpu(), torch.from_numpy(g["train_latents"]), atol=5e-5)
    assert torch.equal(dg.data_test["masks"].cpu(), torch.from_numpy(g["test_masks"]))
    b = dg.get_split("val", 2, chunk=1)
    assert b["measurements"].shape == (15, 2, dg.obs_dim) and b["actions"].shape == (15, 2, 1)
    mb = dg.get_mini_batch("train", 5)
    assert mb["latents"].shape == (15, 5, dg.latent_dim)


@pytest.mark.parametrize("name", CASES)
def test_generator_matches_reference_fixture_hostsim(name):
    subprocess.run(["make", "-C", HS_DIR], check=True, capture_output=True)
    lib = L.HodeLib(HS, required=["hode_abi_version", "hode_last_error", "hode_dopri5_fwd"])
    g, dg = build(name, torch.device("cpu"), lib=lib)
    dg.generate_data()
    check(g, dg)


def test_generator_refuses_cpu_without_the_cuda_library():
    _, dg = build("datagen_d4", torch.device("cpu"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dg.generate_data()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_generator_matches_reference_fixture_gpu(name):
    g, dg = build(name, torch.device("cuda:0"))
    dg.generate_data()
    check(g, dg)


@pytest.mark.gpu
def test_generator_device_rng_large_cohort_statistics():
    """exact_rng=False: per-patient draws on the device.  Distribution checks at 200 k patients, and the latents of the
    first patients must equal an exact_rng-style solve of the same inputs (same kernel, same inputs)."""
    np.random.seed(1)
    N = 200_000
    dg = DataGeneratorRoche(N, 40, 14, 1, RochConfig(kel=1), 0.2, 10, 8, 0.5, p_remove=0.5, output_sparsity=0.625,
                            device=torch.device("cuda:0"), val_size=100, test_size=1000, exact_rng=False)
    dg.generate_data()
    assert torch.isfinite(dg.latents).all()
    assert abs(float(dg.masks.mean()) - 0.5) < 5e-3
    assert abs(float(dg.measurements.mean())) < 1e-3 and abs(float(dg.measurements.std()) - 1.0) < 1e-2
    n_doses = (dg.actions != 0).sum(dim=0).squeeze(-1)
    assert int(n_doses.max()) == 1 and float((n_doses == 1).float().mean()) > 0.999
    assert 4.5 < float(dg.dose_amount.mean()) < 5.5
    again = dg.solve_latents(dg.latents[0], torch.as_tensor(dg.dose_time)[:], torch.as_tensor(dg.dose_amount))
    assert torch.equal(again, dg.latents)


def test_generator_device_rng_path_on_host_emulation():
    """exact_rng=False code path (per-patient draws by torch generators on the cohort device) through the host emulation:
    shapes, dose indexing, mask density, normalisation."""
    subprocess.run(["make", "-C", HS_DIR], check=True, capture_output=True)
    lib = L.HodeLib(HS, required=["hode_abi_version", "hode_last_error", "hode_dopri5_fwd"])
    np.random.seed(3)
    N = 1500
    dg = DataGeneratorRoche(N, 20, 14, 1, RochConfig(kel=1), 0.1, 1, 6, 0.5, p_remove=0.5, output_sparsity=0.5,
                            device=torch.device("cpu"), val_size=100, test_size=200, exact_rng=False, lib=lib)
    dg.generate_data()
    assert dg.latents.shape == (15, N, 6) and torch.isfinite(dg.latents).all()
    assert dg.measurements.shape == (15, N, 20) and dg.masks.shape == (15, N, 20) and dg.actions.shape == (15, N, 1)
    assert abs(float(dg.masks.mean()) - 0.5) < 0.01
    assert torch.allclose(dg.measurements.mean(dim=(0, 1)), torch.zeros(20), atol=1e-4)
    assert torch.allclose(dg.measurements.std(dim=(0, 1)), torch.ones(20), atol=1e-4)
    day = torch.as_tensor(dg.dose_time)[:, 0]
    amt = torch.as_tensor(dg.dose_amount).float()
    assert int(day.min()) >= 0 and int(day.max()) <= 13
    assert torch.equal(dg.actions[day, torch.arange(N), 0], amt)
    assert float((dg.actions != 0).sum()) <= N
    dg.split_sample()
    assert dg.data_train["latents"].shape[1] == N - 300 and dg.data_test["masks"].shape[1] == 200
