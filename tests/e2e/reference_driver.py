"""Run the reference's UNMODIFIED ``experiments/run_simulation.py:run`` (its own ``model.py``, ``training_utils.py``,
``dataloader`` splits) with only the two missing third-party modules supplied: ``torchdiffeq`` and ``properscoring``.

TEST INFRASTRUCTURE.  ``solver='cuda'`` binds ``torchdiffeq.odeint`` to the fused sm_100a kernels
(``hybrid_ode_neurips_2021_b200.install_as_torchdiffeq``): every ``dto(...)`` call of ``model.RocheExpertDecoder.forward``
(``model.py:1116``) and every ``loss.backward()`` of ``training_utils.variational_training_loop`` then runs on the GPU path
with the reference's own vector-field / encoder / VI objects.  ``solver='oracle'`` binds the CPU restatement instead (the
calibration run behind the tolerance band of ``tests/test_reference_e2e.py``).

The cohort is the one ``generated_data/generate_data_train.py`` builds (N = 1 300, D = 6, obs 20, seed 666), produced by the
GPU drop-in generator on CUDA (same random streams bit for bit, latents to 1e-5: ``tests/test_datagen.py``) or by the
reference's own ``lsoda`` generator on the CPU (``generator='reference'``; scipy >= 1.8 needs the copy shim of
``oracle/make_golden_datagen.py``).
"""
from __future__ import annotations

import contextlib
import io
import os
import pickle
import re
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LINE = re.compile(r"Iter (\d+) \| Total Loss ([-\d.eE+naninf]+) \| Train Loss ([-\d.eE+naninf]+)")


def _bind_solver(solver: str):
    from oracle import refload

    for name in ("torchdiffeq",):
        sys.modules.pop(name, None)
    if solver == "cuda":
        import hybrid_ode_neurips_2021_b200 as H

        H.install_as_torchdiffeq(force=True)
        refload.install_shims()  # properscoring only: torchdiffeq is already bound
    else:
        refload.install_shims()
    assert "torchdiffeq" in sys.modules and "properscoring" in sys.modules
    for name in ("model", "training_utils", "experiments.run_simulation"):  # re-import against the solver just bound
        sys.modules.pop(name, None)


def build_cohort(path: str, device: torch.device, generator: str = "dropin"):
    """``generated_data/generate_data_train.py`` with the output path and the device made arguments."""
    from oracle import refload

    sc = refload.load("sim_config")
    dc = sc.DataConfig(n_sample=1300)
    np.random.seed(666)
    torch.manual_seed(666)
    kw = dict(p_remove=dc.p_remove, output_sparsity=0.5, device=device)
    args = (dc.n_sample, dc.obs_dim, dc.t_max, dc.step_size, sc.RochConfig(kel=1), 0.2, 10, dc.latent_dim, dc.sparsity)
    if generator == "reference":
        from oracle.make_golden_datagen import _restore_pinned_scipy_semantics

        _restore_pinned_scipy_semantics()
        dg = refload.load("dataloader").DataGeneratorRoche(*args, **kw)
    else:
        import hybrid_ode_neurips_2021_b200 as H

        dg = H.DataGeneratorRoche(*args, **kw)
    dg.generate_data()
    dg.split_sample()
    dg.set_device(torch.device("cpu"))
    with open(path, "wb") as f:
        pickle.dump(dg, f)
    return dg


def run_hybrid(data_path: str, model_path: str, solver: str, device: str, niters: int, batch_size: int = 10,
               sample: int = 1000, evaluate: bool = True, method: str = "hybrid", elbo: bool = True):
    """``python -m experiments.run_simulation --method=hybrid --batch_size=10 --sample=1000 --restart=1 --arg_itr=niters``
    (the command of ``experiments/Fig3.sh:14``) executed in-process.  Returns the parsed ``Iter`` lines and the stdout."""
    from oracle import refload

    _bind_solver(solver)
    sc = refload.load("sim_config")
    RS = refload.load("experiments.run_simulation")
    TU = refload.load("training_utils")
    if not evaluate:  # CPU calibration only: the 51 x 20 test solves of training_utils.evaluate take ~an hour on the oracle
        RS.training_utils = types.SimpleNamespace(variational_training_loop=TU.variational_training_loop,
                                                  evaluate=lambda *a, **k: None)
    mc = {"hybrid": sc.ModelConfig(path=model_path), "expert": sc.ModelConfig(expert_only=True, path=model_path),
          "neural": sc.ModelConfig(neural_ode=True, path=model_path)}[method]
    oc = sc.OptimConfig(shuffle=False, n_restart=1, batch_size=batch_size, lr=0.01)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        RS.run(666, elbo, device, False, None, data_path, sample, sc.DataConfig(n_sample=sample), sc.RochConfig(), mc, oc,
               sc.EvalConfig(t0=5), arg_itr=niters)
    out = buf.getvalue()
    iters = [(int(m.group(1)), float(m.group(2)), float(m.group(3))) for m in LINE.finditer(out)]
    return iters, out


if __name__ == "__main__":  # calibration: python tests/e2e/reference_driver.py oracle|cuda [niters]
    solver = sys.argv[1] if len(sys.argv) > 1 else "oracle"
    niters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    tmp = os.environ.get("HODE_E2E_TMP", "/tmp/hode_e2e")
    os.makedirs(tmp, exist_ok=True)
    data = os.path.join(tmp, "datafile_dose_exp_{}.pkl".format("ref" if solver == "oracle" else "gpu"))
    if not os.path.exists(data):
        build_cohort(data, torch.device("cuda:0") if solver == "cuda" else torch.device("cpu"),
                     "dropin" if solver == "cuda" else "reference")
    elbo = os.environ.get("HODE_E2E_ELBO", "y") == "y"
    iters, out = run_hybrid(data, os.path.join(tmp, "model_{}_{}".format(solver, int(elbo))) + "/", solver,
                            "0" if solver == "cuda" else "c", niters, evaluate=(solver == "cuda"), elbo=elbo)
    print(out)
    print(iters)
