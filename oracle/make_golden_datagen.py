"""Generate ``tests/golden/datagen_*.npz`` by running the REFERENCE's own ``dataloader.DataGeneratorRoche`` (imported
unmodified from ``/root/reference``; scipy ``lsoda`` per patient, ``dataloader.py:95-266``) exactly as
``generated_data/generate_data_*.py`` do (seed 666 for numpy and torch), at a small cohort size.  Build container only:

    python -m oracle.make_golden_datagen

The fixtures pin the vectorised GPU generator (``hybrid_ode_neurips_2021_b200/datagen.py``): random streams
(coefficients, initial conditions, doses, measurement noise, masks) bit-for-bit, latents to the float32-dopri5 vs
float64-lsoda tolerance.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, DataConfig overrides of generate_data_train.py / generate_data_dim8.py / generate_data_dim12.py)
CASES = {
    "datagen_d6": dict(obs_dim=20, latent_dim=6, output_sparsity=0.5, output_sigma=0.1, dose_max=1),
    "datagen_d8": dict(obs_dim=40, latent_dim=8, output_sparsity=1 - 0.375, output_sigma=0.2, dose_max=10),
    "datagen_d4": dict(obs_dim=20, latent_dim=4, output_sparsity=0.5, output_sigma=0.1, dose_max=10),
}


def _restore_pinned_scipy_semantics():
    import scipy.integrate

    orig = scipy.integrate.ode.integrate
    if getattr(orig, "__hode_copy__", False):
        return

    def integrate(self, t, step=False, relax=False):
        return np.array(orig(self, t, step, relax), copy=True)

    integrate.__hode_copy__ = True
    scipy.integrate.ode.integrate = integrate


def main():
    _restore_pinned_scipy_semantics()
    dl = refload.load("dataloader")
    sc = refload.load("sim_config")
    for name, c in CASES.items():
        n_sample, val_size, test_size = 24, 4, 8
        np.random.seed(666)
        torch.manual_seed(666)
        dg = dl.DataGeneratorRoche(n_sample, c["obs_dim"], 14, 1, sc.RochConfig(kel=1), c["output_sigma"], c["dose_max"],
                                   c["latent_dim"], 0.5, p_remove=0.5, output_sparsity=c["output_sparsity"],
                                   device=torch.device("cpu"), val_size=val_size, test_size=test_size)
        dg.generate_data()
        dg.split_sample()
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            cfg=np.array([n_sample, c["obs_dim"], c["latent_dim"], val_size, test_size], dtype=np.int64),
            cfg_f=np.array([c["output_sigma"], c["dose_max"], 0.5, c["output_sparsity"], 0.5], dtype=np.float64),
            output_coef=dg.output_coef, ml_coef=dg.ml_coef, dose_time=dg.dose_time, dose_amount=dg.dose_amount,
            measurements=dg.measurements.numpy(), actions=dg.actions.numpy(), latents=dg.latents.numpy(),
            masks=dg.masks.numpy(), train_latents=dg.data_train["latents"].numpy(),
            test_masks=dg.data_test["masks"].numpy(),
        )
        print(name, "latents max", float(dg.latents.abs().max()), "train", dg.data_train["latents"].shape)


if __name__ == "__main__":
    main()
