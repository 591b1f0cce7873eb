"""End-to-end level of SURVEY.md section 4: the reference's UNMODIFIED experiment code on the CUDA path.

``experiments/run_simulation.py:run`` -- with the reference's own ``model.py`` (``EncoderLSTM``, ``RocheODE``,
``RocheExpertDecoder``, ``VariationalInference``), ``training_utils.variational_training_loop`` and
``training_utils.evaluate`` -- is executed in-process from the byte-identical copy under ``baseline/_ref`` (made by
``oracle/install_reference.py``; ``/root/reference`` itself does not exist on the GPU box).  The only substitution is the
one the reference leaves open: ``from torchdiffeq import odeint as dto`` (``model.py:10``) binds to
``hybrid_ode_neurips_2021_b200.odeint`` through ``install_as_torchdiffeq()``, so all ~1 500 ``dto`` calls of ten training
iterations + validation + Monte-Carlo evaluation, and every ``loss.backward()``, run the fused sm_100a kernels.

Anchors:
* the published training curve ``results/exp_lhm.csv`` line 1 -- iteration 10: validation total 2486.55, train 241.01 at batch
  10 (``experiments/Fig3.sh:14``).  The curve depends on the random stream of the re-parameterised posterior samples (CPU
  generator in the published run, CUDA generator here), so the band is the one the CPU calibration run of the same driver on
  the restated torchdiffeq shows (``python tests/e2e/reference_driver.py oracle 10``: 2267.62 / 281.58, i.e. -9 % / +17 %);
* a DETERMINISTIC variant (``elbo=False``: the posterior mean is decoded, no sampling): the CUDA path must reproduce the
  value that the same unmodified code gives on the CPU oracle (recorded below from this repository's build container).
"""
import json
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "e2e"))

PUBLISHED_ITER10 = (2486.551712, 241.009094)  # results/exp_lhm.csv:1
# python tests/e2e/reference_driver.py oracle 10 with HODE_E2E_ELBO=n (reference lsoda cohort, CPU oracle solver)
ORACLE_ITER10_NO_ELBO = (1786.919754, 211.086594)


def _ref_available():
    from oracle import refload

    return refload.available()


def test_reference_copy_is_byte_identical():
    """``baseline/_ref`` holds exactly the files of the reference tree (SHA-256 manifest written by the install recipe)."""
    from oracle import install_reference as IR

    man_path = os.path.join(IR.DEST, "MANIFEST.json")
    if not os.path.exists(man_path):
        pytest.skip("baseline/_ref not populated (run __graft_entry__.build() where /root/reference exists)")
    man = json.load(open(man_path))
    assert {"model.py", "training_utils.py", "dataloader.py", "experiments/run_simulation.py", "results/exp_lhm.csv"} <= set(man["files"])
    for rel, dig in man["files"].items():
        assert IR._digest(os.path.join(IR.DEST, rel)) == dig, rel
        src = os.path.join("/root/reference", rel)
        if os.path.exists(src):
            assert IR._digest(src) == dig, rel
    first = open(os.path.join(IR.DEST, "results", "exp_lhm.csv")).readline().strip().split(",")
    assert int(first[0]) == 10 and (float(first[1]), float(first[2])) == PUBLISHED_ITER10


@pytest.fixture(scope="module")
def cohort(tmp_path_factory):
    import reference_driver as RD

    d = tmp_path_factory.mktemp("e2e")
    path = str(d / "datafile_dose_exp.pkl")
    dg = RD.build_cohort(path, torch.device("cuda:0"), "dropin")
    assert tuple(dg.measurements.shape) == (15, 1300, 20) and dg.train_size == 1000
    return path, str(d)


@pytest.mark.gpu
def test_run_simulation_unmodified_reaches_the_published_curve_on_the_cuda_path(cohort):
    if not _ref_available():
        pytest.skip("reference tree not available (baseline/_ref missing)")
    import hybrid_ode_neurips_2021_b200 as H
    import reference_driver as RD

    data, tmp = cohort
    iters, out = RD.run_hybrid(data, os.path.join(tmp, "model_elbo") + "/", "cuda", "0", niters=10, batch_size=10)
    assert sys.modules["torchdiffeq"].odeint is H.odeint and getattr(sys.modules["torchdiffeq"], "__hode_shim__", False)
    assert sys.modules["model"].__file__.endswith(os.path.join("_ref", "model.py")) or "/root/reference" in sys.modules["model"].__file__
    assert H.last_solve_info() is not None and H.last_solve_info().stats is not None  # dopri5 kernels ran
    assert [i[0] for i in iters] == [10], out
    _, total, train = iters[0]
    print("iter 10: total {:.2f} (published {:.2f}), train {:.2f} (published {:.2f})".format(total, PUBLISHED_ITER10[0], train, PUBLISHED_ITER10[1]))
    assert abs(total - PUBLISHED_ITER10[0]) <= 0.20 * PUBLISHED_ITER10[0], (total, out)
    assert abs(train - PUBLISHED_ITER10[1]) <= 0.35 * PUBLISHED_ITER10[1], (train, out)
    # training_utils.evaluate ran on the trained model (1 + 50 decoder solves per test chunk): its four result lines
    for key in ("rmse_z0,", "rmse_x,", "cprs_z0,", "cprs_x,"):
        assert key in out, out
    rmse_x = float([ln for ln in out.splitlines() if ln.startswith("rmse_x,")][0].split(",")[1])
    assert 0.3 < rmse_x < 3.0  # ten iterations in: finite and of the order of the normalised measurements
    assert os.path.exists(os.path.join(tmp, "model_elbo", "VI_LSTMEncoder_HybridDecoder.pkl"))  # the reference's checkpoint name


@pytest.mark.gpu
def test_deterministic_training_matches_the_cpu_oracle_run(cohort):
    if not _ref_available():
        pytest.skip("reference tree not available (baseline/_ref missing)")
    import reference_driver as RD

    data, tmp = cohort
    iters, out = RD.run_hybrid(data, os.path.join(tmp, "model_noelbo") + "/", "cuda", "0", niters=10, batch_size=10,
                               elbo=False, evaluate=False)
    assert [i[0] for i in iters] == [10], out
    _, total, train = iters[0]
    ref_total, ref_train = ORACLE_ITER10_NO_ELBO
    print("iter 10 (elbo=False): total {:.4f} vs CPU oracle {:.4f}; train {:.4f} vs {:.4f}".format(total, ref_total, train, ref_train))
    # ten Adam steps through ~350 accepted dopri5 steps each, float32 on both sides, cohort from the float32 GPU generator vs
    # float64 lsoda (measurements agree to 3e-5).  Observed on B200: 1.7e-6 / 2e-6 relative; the gate leaves room for the
    # launch-to-launch reordering of the atomically reduced parameter gradients.
    assert abs(total - ref_total) <= 2e-4 * ref_total, (total, ref_total)
    assert abs(train - ref_train) <= 2e-4 * ref_train, (train, ref_train)
