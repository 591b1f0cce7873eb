"""Import the reference's own ``model.py`` UNMODIFIED, with the two missing third-party modules injected.

TEST INFRASTRUCTURE ONLY.  The tree is ``/root/reference`` where that exists (the build container) and otherwise the
byte-identical copy ``baseline/_ref`` made by ``oracle/install_reference.py`` (git-ignored; it travels to the GPU box).
``model.py:10`` needs ``torchdiffeq`` and ``training_utils.py:4`` needs ``properscoring``; neither is installable
offline, so ``oracle.odeint`` stands in for the former and a minimal ``crps_ensemble`` for the latter.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("HODE_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "baseline", "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isfile(os.path.join(c, "model.py"))), "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def _crps_ensemble(observations, forecasts):
    """CRPS of an ensemble forecast, ``E|X - y| - 0.5 E|X - X'|`` (what properscoring.crps_ensemble computes)."""
    import numpy as np

    obs = np.asarray(observations, dtype=float)
    fc = np.asarray(forecasts, dtype=float)
    if fc.ndim == obs.ndim:
        fc = fc[..., None]
    term1 = np.abs(fc - obs[..., None]).mean(axis=-1)
    term2 = np.abs(fc[..., :, None] - fc[..., None, :]).mean(axis=(-2, -1))
    return term1 - 0.5 * term2


def install_shims():
    from . import odeint as oi

    if "torchdiffeq" not in sys.modules:
        m = types.ModuleType("torchdiffeq")
        m.odeint = oi.odeint
        m.__oracle_shim__ = True
        sys.modules["torchdiffeq"] = m
    if "properscoring" not in sys.modules:
        p = types.ModuleType("properscoring")
        p.crps_ensemble = _crps_ensemble
        p.__oracle_shim__ = True
        sys.modules["properscoring"] = p


def load(name: str = "model"):
    """Return the reference module ``name`` (``model``, ``training_utils``, ``dataloader``, ``sim_config`` ...)."""
    if not available():
        raise FileNotFoundError("reference tree not present at " + REFERENCE_ROOT)
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module(name)
