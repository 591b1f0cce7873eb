"""hybrid_ode_neurips_2021_b200 -- B200 (sm_100a) implementation of the hybrid-ODE hot path of
ZhaozhiQIAN/Hybrid-ODE-NeurIPS-2021: fused fixed-step / dopri5 integration (forward + discrete backprop, and the
tape-free continuous adjoint for the fixed-step methods) of the expert PK/PD + latent-MLP vector field, the fused
read-out + masked-SSE reduction, one-launch Monte-Carlo evaluation with CRPS, the real-data fields, and the cohort
generator that feeds the path.

Importing this package does not load CUDA; the shared library is loaded on first solver call and its absence is a
hard error (there is no CPU fallback).
"""
from .solver import odeint, odeint_adjoint, odeint_ensemble, odeint_sse, EnsembleParams, last_solve_info, last_adjoint_solve_info, fixed_grid_points  # noqa: F401
from .model import RocheODE, NeuralODE, RocheExpertDecoder, RochConfig  # noqa: F401
from .real import RocheODEReal, NeuralODEReal, NeuralODEReal2nd, DecoderReal  # noqa: F401
from .loss import masked_sse, decode_sse_loss  # noqa: F401
from .integrate import install_as_torchdiffeq, patch_model  # noqa: F401
from .datagen import DataGeneratorRoche  # noqa: F401
from .evaluation import (crps_ensemble, mc_solve, decode_crps, evaluate_chunk, evaluate, evaluate_horizon,  # noqa: F401
                         evaluate_ensemble, evaluate_ensemble_horizon, bootstrap_RMSE)

__version__ = "0.1.0"
