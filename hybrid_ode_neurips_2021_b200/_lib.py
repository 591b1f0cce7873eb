"""ctypes binding of ``libhode_b200.so`` (the C ABI declared in ``include/hode.h``).

There is no fallback: if the shared library is missing or does not export a symbol, importing a solver raises.
The library is built in-tree by ``hybrid_ode_neurips_2021_b200/build.py`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libhode_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)

# enums of include/hode.h
FIELD_ROCHE, FIELD_NEURAL = 0, 1
FIELD_ROCHE_REAL, FIELD_NEURAL_REAL, FIELD_NEURAL_REAL_2ND = 2, 3, 4
EULER, MIDPOINT, RK4_38, DOPRI5 = 0, 1, 2, 3
CTRL_BATCH, CTRL_TRAJ = 0, 1
FLAG_HILL2, FLAG_ABLATE, FLAG_ADJ_SEMINORM = 1, 2, 4
METHODS = {"euler": EULER, "midpoint": MIDPOINT, "rk4": RK4_38, "dopri5": DOPRI5}
SOLVE_OK, SOLVE_DT_UNDERFLOW, SOLVE_NONFINITE, SOLVE_MAX_STEPS, SOLVE_TAPE_FULL = 0, 1, 2, 3, 4
OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = 0, -1, -2, -3, -4


class HodeCfg(C.Structure):
    _fields_ = [
        ("field", C.c_int32),
        ("latent_dim", C.c_int32),
        ("method", C.c_int32),
        ("controller", C.c_int32),
        ("perturb", C.c_int32),
        ("n_dose", C.c_int32),
        ("expert_grads", C.c_int32),
        ("flags", C.c_int32),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("safety", C.c_double),
        ("ifactor", C.c_double),
        ("dfactor", C.c_double),
        ("first_step", C.c_double),
        ("max_num_steps", C.c_int64),
        ("attempt_cap", C.c_int64),
    ]


_P = C.c_void_p
_I32, _I64, _F64 = C.c_int32, C.c_int64, C.c_double
_CFG = C.POINTER(HodeCfg)

# name -> (restype, argtypes); every symbol include/hode.h declares
SIGNATURES = {
    "hode_abi_version": (_I32, []),
    "hode_last_error": (C.c_char_p, []),
    "hode_param_count": (_I64, [_CFG]),
    "hode_supported": (_I32, [_CFG]),
    "hode_dose_schedule": (_I32, [_P, _I64, _I64, _I32, _I64, _P, _P, _P, _P]),
    "hode_fixed_tape_bytes": (C.c_size_t, [_CFG, _I64, _I32]),
    "hode_fixed_fwd": (_I32, [_CFG, _I64, _I64, _P, _P, _P, _I64, _P, _P, _P, _I32, _P, _I32, _P, _P, _P]),
    "hode_fixed_fwd_sse_supported": (_I32, [_CFG, _I32, _I32]),
    "hode_fixed_fwd_sse": (_I32, [_CFG, _I64, _P, _P, _P, _I64, _P, _P, _I32, _P, _I32, _P, _P, _I32, _P, _P, _F64, _P, _P, _P,
                                  _P, _P, _P, _P]),
    "hode_fixed_bwd": (_I32, [_CFG, _I64, _I64, _P, _P, _I64, _P, _P, _I32, _P, _I32, _P, _I32, _P, _P, _P, _P, _P]),
    "hode_fixed_adjoint": (_I32, [_CFG, _I64, _I64, _P, _P, _I64, _P, _P, _I32, _P, _I32, _P, _I32, _P, _P, _P, _P, _P]),
    "hode_dopri5_max_batch": (_I64, [_CFG]),
    "hode_dopri5_fwd": (_I32, [_CFG, _I64, _I64, _P, _P, _P, _I64, _P, _P, _P, _I32, _P, _P, _P, _I32, _P, _P]),
    "hode_dopri5_bwd": (
        _I32,
        [_CFG, _I64, _I64, _P, _P, _I64, _P, _P, _I32, _P, _I32, _P, _P, _P, _I32, _P, _P, _P, _P],
    ),
    "hode_dopri5_adjoint": (_I32, [_CFG, _I64, _I64, _P, _P, _I64, _P, _P, _I32, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "hode_bench_ffma": (_I64, [_I32, _I32, _P, _P]),
    "hode_real_param_count": (_I64, [_I32, _I32, _I32]),
    "hode_real_dose_tables": (_I32, [_I32, _P, _I64, _I64, _I32, _I64, _P, _P, _P]),
    "hode_real_fixed_fwd": (_I32, [_I32, _I32, _I32, _I32, _I32, _I64, _P, _P, _I32, _P, _P, _I32, _P, _I32, _P, _P, _P]),
    "hode_real_fixed_bwd": (_I32, [_I32, _I32, _I32, _I32, _I32, _I64, _P, _I32, _P, _P, _I32, _P, _I32, _P, _P, _P, _P, _P]),
    "hode_crps_ensemble": (_I32, [_P, _P, _I64, _I32, _I64, _I64, _P, _P]),
    "hode_decode_crps": (_I32, [_I32, _I32, _I32, _I64, _I32, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P]),
    "hode_decode_sse": (_I32, [_I32, _I32, _I32, _I64, _F64, _P, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P, _P, _P, _P]),
}


class HodeError(RuntimeError):
    pass


class HodeLib:
    """A loaded ``libhode`` with typed entry points.  ``required`` lists the symbols that must resolve."""

    def __init__(self, path: str = LIB_PATH, required=None):
        if not os.path.isfile(path):
            raise HodeError(
                "{} not found: the CUDA extension has not been built (run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `python -m hybrid_ode_neurips_2021_b200.build`). There is no CPU fallback.".format(path)
            )
        self.path = path
        self.dll = C.CDLL(path)
        names = list(SIGNATURES) if required is None else list(required)
        for name in names:
            try:
                fn = getattr(self.dll, name)
            except AttributeError as e:
                raise HodeError("{} does not export {}".format(path, name)) from e
            fn.restype, fn.argtypes = SIGNATURES[name]
            setattr(self, name, fn)
        if self.hode_abi_version() != 1:
            raise HodeError("ABI version mismatch in " + path)

    def check(self, rc: int, what: str):
        if rc == OK:
            return
        msg = self.hode_last_error()
        msg = msg.decode() if msg else ""
        if rc == ERR_UNSUPPORTED:
            raise NotImplementedError("{}: {}".format(what, msg))
        if rc == ERR_ARG:
            raise ValueError("{}: {}".format(what, msg))
        raise HodeError("{}: {} (rc={})".format(what, msg, rc))


_lock = threading.Lock()
_lib = None


def get_lib() -> HodeLib:
    """The process-wide CUDA library (loaded on first use)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                _lib = HodeLib(LIB_PATH)
    return _lib
