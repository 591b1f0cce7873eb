"""Hooks for running the reference's UNMODIFIED experiment code on top of this package.

``install_as_torchdiffeq()`` registers a module named ``torchdiffeq`` whose ``odeint`` is the fused solver, so that
``from torchdiffeq import odeint as dto`` (``/root/reference/model.py:10``) binds to it: the reference's own
``RocheODE`` / ``NeuralODE`` instances are recognised by structure and integrated by the sm_100a kernels
(``odeint_adjoint``, the import the reference keeps commented out at ``model.py:9``, is provided as well).
``patch_model(module)`` additionally swaps the hot-path classes of an imported reference ``model`` module for the drop-ins
(vectorised ``set_action``, one-launch solves).  The callers either side of the path -- ``EncoderLSTM``,
``VariationalInference``, ``training_utils`` -- stay the reference's own, unmodified code.
"""
from __future__ import annotations

import sys
import types

from . import model as _model
from .solver import odeint, odeint_adjoint


def install_as_torchdiffeq(force: bool = False):
    if "torchdiffeq" in sys.modules and not force and not getattr(sys.modules["torchdiffeq"], "__hode_shim__", False):
        raise RuntimeError("a real torchdiffeq is already imported; pass force=True to shadow it")
    m = types.ModuleType("torchdiffeq")
    m.odeint = odeint
    m.odeint_adjoint = odeint_adjoint  # the alternative import the reference keeps commented out (model.py:9)
    m.__hode_shim__ = True
    sys.modules["torchdiffeq"] = m
    return m


def patch_model(module):
    module.dto = odeint
    module.RocheODE = _model.RocheODE
    module.NeuralODE = _model.NeuralODE
    module.RocheExpertDecoder = _model.RocheExpertDecoder
    from . import real as _real

    module.RocheODEReal = _real.RocheODEReal
    module.NeuralODEReal = _real.NeuralODEReal
    module.NeuralODEReal2nd = _real.NeuralODEReal2nd
    module.DecoderReal = _real.DecoderReal
    return module
