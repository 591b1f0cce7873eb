"""CPU-side check of the real-data field SOURCE (csrc/hode_real.cuh: RocheODEReal / NeuralODEReal / NeuralODEReal2nd, their
hand-derived VJPs and the O(1) dose tables) through the test-only host emulation, against the oracle restatement of
model.py:570-862 (which tests/test_oracle_fields.py pins bit-for-bit to the reference's own classes)."""
import os
import subprocess

import pytest
import torch

from hybrid_ode_neurips_2021_b200 import _lib as L
from hybrid_ode_neurips_2021_b200 import ops
from hybrid_ode_neurips_2021_b200 import real as R
from oracle import fields as OF
from oracle import odeint as OI

from _util import relerr

HS_DIR = os.path.join(os.path.dirname(__file__), "hostsim")
HS = os.path.join(HS_DIR, "libhode_hostsim.so")
SYMS = ["hode_abi_version", "hode_last_error", "hode_real_param_count", "hode_real_dose_tables", "hode_real_fixed_fwd",
        "hode_real_fixed_bwd"]


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-C", HS_DIR], check=True, capture_output=True)
    return L.HodeLib(HS, required=SYMS)


def icu_cohort(B, Z, T, seed):
    g = torch.Generator().manual_seed(seed)
    y0 = torch.randn(B, Z, generator=g) * 0.3
    a = torch.rand(T, B, 1, generator=g) * (torch.rand(T, B, 1, generator=g) < 0.25).float()
    s = torch.randn(T, B, 11, generator=g)
    return y0, a, s


def make_oracle(kind, Z, H, seed):
    torch.manual_seed(seed)
    if kind == L.FIELD_ROCHE_REAL:
        o = OF.OracleRocheODEReal(Z, H)
        with torch.no_grad():
            o.kel.fill_(0.23); o.kel2.fill_(0.17); o.k_immunity.fill_(0.9)
    else:
        o = OF.OracleNeuralODEReal(Z, H, second=(kind == L.FIELD_NEURAL_REAL_2ND))
    o.static_dim = 11  # attribute the structural recogniser looks at
    return o


CASES = [(L.FIELD_ROCHE_REAL, 20, 43), (L.FIELD_ROCHE_REAL, 4, 43), (L.FIELD_NEURAL_REAL, 20, 43),
         (L.FIELD_NEURAL_REAL_2ND, 40, 43), (L.FIELD_NEURAL_REAL, 4, 7), (L.FIELD_NEURAL_REAL_2ND, 8, 64)]


@pytest.mark.parametrize("kind,Z,H", CASES)
@pytest.mark.parametrize("method", ["midpoint", "rk4", "euler"])
def test_real_fields_forward_and_reverse_sweep(lib, kind, Z, H, method):
    B, T, t0 = 4, 40, 24
    o = make_oracle(kind, Z, H, seed=Z + H)
    y0, a, s = icu_cohort(B, Z, T, seed=Z)
    o.set_action_static(a, s)
    t = torch.arange(t0 - 1, T, 1.0)
    opts = {"step_size": 1.0, "perturb": True}
    W = torch.randn(len(t), B, Z, generator=torch.Generator().manual_seed(1))
    z = y0.clone().requires_grad_(True)
    ref = OI.odeint(o, z, t, method=method, options=opts)
    (ref * W).sum().backward()
    assert R.real_field_kind(o) == kind
    params = R.pack_real_params(o, kind).detach().contiguous()
    assert params.numel() == lib.hode_real_param_count(kind, Z, H)
    tab = ops.real_dose_tables(lib, kind, a, params)
    grid = OI.fixed_grid_points(t, 1.0).contiguous()
    h, tape = ops.real_fixed_fwd(lib, kind, Z, H, L.METHODS[method], True, y0, tab, params, grid, t, True)
    gy0, gp = ops.real_fixed_bwd(lib, kind, Z, H, L.METHODS[method], True, tab, params, grid, t, W, tape)
    assert relerr(h, ref) < 5e-6
    assert relerr(gy0, z.grad) < 2e-5
    gref = torch.cat([p.grad.reshape(-1) for p in R.pack_real_params_list(o, kind)])
    assert relerr(gp, gref) < 5e-5


def test_dose_tables_equal_the_reference_sums(lib):
    """O(1) table look-up == the reference's O(T) sums (model.py:653-657, 753-760), including perturbed stage times."""
    B, T = 6, 50
    _, a, s = icu_cohort(B, 20, T, seed=2)
    o = make_oracle(L.FIELD_ROCHE_REAL, 20, 43, seed=1)
    o.set_action_static(a, s)
    params = R.pack_real_params(o, L.FIELD_ROCHE_REAL).detach().contiguous()
    tab = ops.real_dose_tables(lib, L.FIELD_ROCHE_REAL, a, params)
    kel = float(o.kel.detach())
    for t in (0.0, 0.5, 1.0, 23.0, 24.000002, 24.5, 48.999996, 49.0, 50.0, 57.25):
        n = min(T, int(torch.floor(torch.tensor(t))))
        got = torch.exp(torch.tensor(kel * (n - t))) * tab[0, n]
        ref = o.dose_at_time(torch.tensor(t)).detach()
        assert relerr(got, ref) < 2e-6 or float(ref.abs().max()) == 0.0, t
    n_ = OF.OracleNeuralODEReal(20, 43)
    n_.set_action_static(a, s)
    tabn = ops.real_dose_tables(lib, L.FIELD_NEURAL_REAL, a, params)
    for t in (0.0, 0.99, 23.0, 24.000002, 48.999996, 49.0, 50.0, 61.0):
        n = int(t)
        got = tabn[0, n] if n < T else torch.zeros(B)
        # torch.cumsum adds in scan order, the table sequentially: last-bit differences only
        assert torch.allclose(got, n_.dose_at_time(torch.tensor(t))[:, 0], rtol=2e-6, atol=0), t


def test_neural_real_negative_time_index_wraps_like_python(lib):
    """``cumsum(action)[int(t)]`` (model.py:753-760) with ``int(t) == -1`` is Python negative indexing: the LAST row.  It is
    what a direct ``odeint`` call sees at ``t[0] = -1`` without ``perturb`` (DecoderReal's default grid starts at ``t0 - 1``)."""
    B, T, Z, H = 3, 12, 4, 7
    kind = L.FIELD_NEURAL_REAL
    o = make_oracle(kind, Z, H, seed=3)
    y0, a, s = icu_cohort(B, Z, T, seed=5)
    o.set_action_static(a, s)
    t = torch.arange(-1.0, 4.0, 1.0)
    ref = OI.odeint(o, y0, t, method="euler", options={"step_size": 1.0})
    assert float((o.dose_at_time(torch.tensor(-1.0)) - torch.cumsum(a, 0)[-1]).abs().max()) == 0.0
    params = R.pack_real_params(o, kind).detach().contiguous()
    tab = ops.real_dose_tables(lib, kind, a, params)
    grid = OI.fixed_grid_points(t, 1.0).contiguous()
    h, _ = ops.real_fixed_fwd(lib, kind, Z, H, L.EULER, False, y0, tab, params, grid, t, False)
    assert relerr(h, ref) < 5e-6
