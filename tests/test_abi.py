"""The C-ABI library loads and exports every symbol include/hode.h declares (no compute calls: no GPU here), and the
host-side layer fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

import hybrid_ode_neurips_2021_b200 as H
from hybrid_ode_neurips_2021_b200 import _lib as L
from hybrid_ode_neurips_2021_b200 import ops, solver
from oracle import odeint as OI

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "hode.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hode_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = header_functions()
    assert set(names) == set(L.SIGNATURES), (names, sorted(L.SIGNATURES))
    lib = L.get_lib()
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.hode_abi_version() == 1


def test_cfg_struct_layout_and_param_counts():
    assert ctypes.sizeof(L.HodeCfg) == 8 * 4 + 6 * 8 + 2 * 8
    lib = L.get_lib()
    for D, p in ((4, 13), (6, 27), (8, 49), (12, 117)):
        cfg = ops.make_cfg(L.FIELD_ROCHE, D, L.DOPRI5)
        assert lib.hode_param_count(ctypes.byref(cfg)) == p
        assert lib.hode_supported(ctypes.byref(cfg)) == 1
        m = H.RocheODE(D, 1, 14, 1, device="cpu")
        assert solver.pack_params(m, L.FIELD_ROCHE).numel() == p
    for D, p in ((6, 847), (8, 1449), (12, 3133)):
        cfg = ops.make_cfg(L.FIELD_NEURAL, D, L.RK4_38)
        assert lib.hode_param_count(ctypes.byref(cfg)) == p
        m = H.NeuralODE(D, 1, 14, 1, device="cpu")
        assert solver.pack_params(m, L.FIELD_NEURAL).numel() == p
    cfg = ops.make_cfg(L.FIELD_ROCHE, 5, L.DOPRI5)
    assert lib.hode_supported(ctypes.byref(cfg)) == 0


def test_ensemble_params_rehomes_members_on_one_flat_leaf():
    torch.manual_seed(3)
    for make, kind, D in ((H.RocheODE, L.FIELD_ROCHE, 8), (H.NeuralODE, L.FIELD_NEURAL, 6)):
        members = [make(D, 1, 14, 1, device="cpu") for _ in range(3)]
        packed = torch.stack([solver.pack_params(m, kind) for m in members]).detach().clone()
        sds = [{k: v.clone() for k, v in m.state_dict().items()} for m in members]
        ens = H.EnsembleParams(members)
        assert len(ens) == 3 and ens.parameters()[0] is ens.flat and ens.flat.requires_grad
        assert torch.equal(ens.flat.detach(), packed)  # rows are the packed layout of include/hode.h
        for m, sd in zip(members, sds):
            assert list(m.state_dict().keys()) == list(sd.keys())
            assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
            assert torch.equal(solver.pack_params(m, kind).detach(), packed[members.index(m)])
        with torch.no_grad():
            ens.flat.mul_(2.0)  # an optimizer step on the flat tensor is seen by the members
        assert torch.equal(solver.pack_params(members[1], kind).detach(), 2.0 * packed[1])
        members[0].load_state_dict(sds[0])  # and a member's checkpoint load writes through
        assert torch.equal(ens.flat[0].detach(), packed[0])
    assert ens.hill_exponents_are_two() is False  # NeuralODE members: no Hill flag
    r = H.EnsembleParams([H.RocheODE(6, 1, 14, 1, device="cpu") for _ in range(2)])
    assert r.hill_exponents_are_two() is True
    with torch.no_grad():
        r.flat[1, 0] = 1.5
    assert r.hill_exponents_are_two() is False  # cached per version of the flat tensor
    with pytest.raises(ValueError):
        H.EnsembleParams([H.RocheODE(6, 1, 14, 1, device="cpu"), H.RocheODE(8, 1, 14, 1, device="cpu")])


def test_argument_errors_are_reported_not_crashed():
    lib = L.get_lib()
    cfg = ops.make_cfg(L.FIELD_ROCHE, 6, L.RK4_38)
    rc = lib.hode_fixed_fwd(ctypes.byref(cfg), -1, 1, None, None, None, 1, None, None, None, 1, None, 1, None, None, None)
    assert rc == L.ERR_ARG and b"negative" in lib.hode_last_error()
    rc = lib.hode_dopri5_fwd(ctypes.byref(cfg), 1, 1, None, None, None, 1, None, None, None, 1, None, None, None, 0, None, None)
    assert rc == L.ERR_ARG
    with pytest.raises(ValueError):
        lib.check(rc, "x")


def test_no_cpu_fallback_and_type_errors():
    m = H.RocheODE(6, 1, 14, 1, device="cpu")
    a = torch.zeros(15, 2, 1)
    a[3, :, 0] = 1.0
    m.set_action(a)
    assert m.times.tolist() == [[3], [3]] and m.dosage.tolist() == [1.0, 1.0]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.odeint(m, torch.zeros(2, 6), torch.arange(0, 15.0))
    with pytest.raises(TypeError):
        H.odeint(torch.nn.Linear(6, 6), torch.zeros(2, 6), torch.arange(0, 15.0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.decode_sse_loss(torch.zeros(15, 2, 6), torch.zeros(20, 6), torch.zeros(20), torch.zeros(15, 2, 20), torch.zeros(15, 2, 20))
    with pytest.raises(L.HodeError, match="not found"):
        L.HodeLib("/nonexistent/libhode_b200.so")


def test_drop_in_surface():
    dec = H.RocheExpertDecoder(20, 6, 1, 14, 1, device="cpu")
    assert dec.model_name == "HybridDecoder" and dec.t.tolist() == list(range(15))
    assert H.RocheExpertDecoder(20, 4, 1, 14, 1, device="cpu").model_name == "ExpertDecoder"
    nd = H.RocheExpertDecoder(20, 6, 1, 14, 1, roche=False, device="cpu")
    assert nd.model_name == "NeuralODEDecoder"
    assert list(nd.state_dict().keys()) == ["output_function.0.weight", "output_function.0.bias", "ode.kel",
                                            "ode.ml_net.0.weight", "ode.ml_net.0.bias", "ode.ml_net.2.weight", "ode.ml_net.2.bias"]
    assert list(dec.state_dict().keys())[2:15] == ["ode." + n for n in solver.EXPERT_NAMES]
    # the eager field stays callable and equals the oracle restatement
    from oracle import fields as OF
    o = OF.OracleRocheODE(6)
    o.load_state_dict(dec.ode.state_dict())
    a = torch.zeros(15, 3, 1)
    a[2, :, 0] = 4.0
    o.set_action(a)
    dec.ode.set_action(a)
    y = torch.rand(3, 6)
    assert torch.equal(o(torch.tensor(3.5), y), dec.ode(torch.tensor(3.5), y))


def test_fixed_grid_points_equal_the_oracle():
    for t, h in ((torch.arange(0, 15.0), 0.0625), (torch.arange(0, 15.0), 0.3), (torch.tensor([1.0, 2.5, 7.0]), 0.05),
                 (torch.arange(0, 15.0), None)):
        assert torch.equal(solver.fixed_grid_points(t, h), OI.fixed_grid_points(t, h))


def test_adjoint_surface_and_reversed_time_grids():
    """odeint_adjoint: torchdiffeq's signature, dopri5 with the mixed norm (where built) or the seminorm, no CPU fallback; the adjoint's per-interval grids
    are torchdiffeq's fixed grid of the negated interval (what the oracle's reverse-time odeint constructs)."""
    m = H.RocheODE(6, 1, 14, 1, device="cpu")
    a = torch.zeros(15, 2, 1)
    a[3, :, 0] = 1.0
    m.set_action(a)
    t = torch.arange(0, 15.0)
    with pytest.raises(NotImplementedError, match="seminorm"):  # the default mixed norm needs the batch-coupled controller ...
        H.odeint_adjoint(m, torch.zeros(2, 6), t, options={"controller": "trajectory"})
    with pytest.raises(NotImplementedError, match="seminorm"):  # ... and RocheODE up to latent_dim 8
        H.odeint_adjoint(H.NeuralODE(6, 1, 14, 1, device="cpu"), torch.zeros(2, 6), t)
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # default method dopri5, default (mixed) adjoint norm
        H.odeint_adjoint(m, torch.zeros(2, 6), t)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.odeint_adjoint(m, torch.zeros(2, 6), t, adjoint_options={"norm": "seminorm"})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.odeint_adjoint(m, torch.zeros(2, 6), t, method="rk4", options={"step_size": 0.25})
    with pytest.raises(ValueError, match="Invalid method"):
        H.odeint_adjoint(m, torch.zeros(2, 6), t, method="rk45")
    with pytest.raises(NotImplementedError, match="adjoint_params"):
        H.odeint_adjoint(m, torch.zeros(2, 6), t, method="rk4", adjoint_params=(torch.nn.Parameter(torch.zeros(1)),))
    for tt, h in ((t, 0.0625), (t, 0.3), (torch.tensor([1.0, 2.5, 7.0]), 0.05), (t, None)):
        grid, counts = solver.adjoint_grid_points(tt, h)
        assert counts.dtype == torch.int32 and counts.numel() == tt.numel() - 1 and int(counts.sum()) == grid.numel()
        off = 0
        for iv, i in enumerate(range(tt.numel() - 1, 0, -1)):
            seg = grid[off:off + int(counts[iv])]
            assert torch.equal(seg, OI.fixed_grid_points(-(tt[i - 1:i + 1].flip(0)), h))
            assert seg[0] == -tt[i] and seg[-1] == -tt[i - 1] and bool((seg[1:] > seg[:-1]).all())
            off += int(counts[iv])
    import sys, types
    saved = sys.modules.get("torchdiffeq")
    try:
        sys.modules.pop("torchdiffeq", None)
        shim = H.install_as_torchdiffeq()
        assert shim.odeint is H.odeint and shim.odeint_adjoint is H.odeint_adjoint
    finally:
        if saved is not None:
            sys.modules["torchdiffeq"] = saved
        else:
            sys.modules.pop("torchdiffeq", None)
