/*
 * hode.h -- C ABI of libhode_b200.so: the B200 (sm_100a) implementation of the hybrid-ODE hot path of
 * ZhaozhiQIAN/Hybrid-ODE-NeurIPS-2021.
 *
 * The reference is pure Python; the interface each entry point replaces is cited as reference file:line
 * (paths relative to the reference root; "torchdiffeq" = the un-vendored torchdiffeq==0.2.2 of requirements.txt:9).
 * The Python binding a maintainer adds on the reference side is a ctypes stub, shown in INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer on the current CUDA device unless named host_*; the library never
 *    allocates, frees or retains memory; work is enqueued on `stream` (a cudaStream_t passed as void*) and the call
 *    returns without synchronising.  Thread-safe.  The only library state is one per-device constant-memory copy of the
 *    staged parameters used by launches with a single parameter set (param_set_of_group == NULL): such launches issued on
 *    DIFFERENT streams of one device are ordered by an internal event (they would each fill the GPU anyway);
 *  - return value: HODE_OK, or a negative hode_status; hode_last_error() gives a per-thread message;
 *    solver failures (torchdiffeq's assertions) are reported per controller group in `stats`, not by return value;
 *  - trajectories are independent patients.  n_traj = n_groups * batch.  A "group" is what ONE reference odeint
 *    call integrates (model.py:1116): it shares a parameter set and -- for the batch-coupled dopri5 controller --
 *    one step-size sequence.  Trajectory index = group * batch + b.
 *  - layouts are the reference's own: y0 [n_traj, D] row-major (model.py:1112 `init`), solution h
 *    [n_t, n_traj, D] time-major (what torchdiffeq.odeint returns), observations x/mask [n_t, n_traj, obs]
 *    (model.py:1151-1153) with arbitrary element strides.
 *  - packed parameter vector of one parameter set (float32, length hode_param_count()):
 *      ROCHE : 13 expert scalars in named_parameters() order (model.py:468-481: HillCure, HillPatho, ec50_patho,
 *              emax_patho, k_dexa, k_discure_immunereact, k_discure_immunity, k_disprog, k_immune_disease,
 *              k_immune_feedback, k_immune_off, k_immunity, kel), then ml_net[0].weight [D-4, D] row-major, then
 *              ml_net[0].bias [D-4]  (model.py:488)
 *      NEURAL: kel, ml_net[0].weight [10D, D+1], ml_net[0].bias [10D], ml_net[2].weight [D, 10D],
 *              ml_net[2].bias [D]  (model.py:989-996)
 */
#ifndef HODE_H_
#define HODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HODE_ABI_VERSION 1

typedef enum hode_status {
    HODE_OK = 0,
    HODE_ERR_ARG = -1,         /* bad argument (null pointer, negative size, ...) */
    HODE_ERR_UNSUPPORTED = -2, /* latent_dim / method / field combination not compiled in */
    HODE_ERR_CUDA = -3,        /* CUDA launch or runtime error */
    HODE_ERR_NO_DEVICE = -4    /* no sm_100 device / no kernel image for this device */
} hode_status;

/* vector field: model.py:446-555 (RocheODE, ablate=False) / model.py:969-1026 (NeuralODE) */
typedef enum hode_field {
    HODE_FIELD_ROCHE = 0,
    HODE_FIELD_NEURAL = 1,
    /* real-data (ICU) fields, hode_real_* entry points only: model.py:570-657, 717-769, 660-714 */
    HODE_FIELD_ROCHE_REAL = 2,
    HODE_FIELD_NEURAL_REAL = 3,
    HODE_FIELD_NEURAL_REAL_2ND = 4
} hode_field;

/* torchdiffeq SOLVERS keys used by the reference: 'euler', 'midpoint', 'rk4' (= 3/8 rule), 'dopri5' */
typedef enum hode_method { HODE_EULER = 0, HODE_MIDPOINT = 1, HODE_RK4_38 = 2, HODE_DOPRI5 = 3 } hode_method;

/* dopri5 step-size control.  BATCH = torchdiffeq semantics: one RMS error norm, one dt, one accept/reject per
 * group (per odeint call).  TRAJ = every trajectory is its own controller (== B calls of batch 1). */
typedef enum hode_controller { HODE_CTRL_BATCH = 0, HODE_CTRL_TRAJ = 1 } hode_controller;

/* per-group solver status (stats[g].status); 1..3 are torchdiffeq's assertion failures */
typedef enum hode_solve_status {
    HODE_SOLVE_OK = 0,
    HODE_SOLVE_DT_UNDERFLOW = 1, /* 'underflow in dt {}' */
    HODE_SOLVE_NONFINITE = 2,    /* 'non-finite values in state `y`: {}' */
    HODE_SOLVE_MAX_STEPS = 3,    /* 'max_num_steps exceeded ({}>={})' */
    HODE_SOLVE_TAPE_FULL = 4     /* accepted steps exceeded tape_capacity (retry with a larger tape) */
} hode_solve_status;

typedef struct hode_stats {
    int32_t accepted; /* accepted steps (== tape length) */
    int32_t rejected; /* rejected attempts */
    int32_t nfe;      /* vector-field evaluations, torchdiffeq accounting: 2 + 6*(accepted+rejected) */
    int32_t status;   /* hode_solve_status */
} hode_stats;

/* hode_cfg.flags.  HODE_FLAG_HILL2: the caller guarantees HillCure == HillPatho == 2.0 exactly (the RochConfig
 * defaults, sim_config.py:5-6; never trained by the simulation experiments) in EVERY parameter set of the call; the
 * library then runs kernels in which x**Hill is a multiply (model.py:529, 537-538).  A violated guarantee is detected on
 * the device and reported loudly: NaN in the solution / gradients (dopri5: status HODE_SOLVE_NONFINITE).
 * HODE_FLAG_ABLATE: RocheODE(ablate=True), the ablation study's expert part (model.py:545-549):
 * dx = (ImmuneReact, -Disease theta_1, Dose2, -Immunity theta_2); the packed parameters gain theta_1, theta_2 at the END
 * (after ml_net's bias); the dose schedule is unused.
 * HODE_FLAG_ADJ_SEMINORM (hode_dopri5_adjoint): the adjoint solve is controlled by torchdiffeq's 'seminorm' (state and state
 * adjoint; the parameter adjoints are integrated but take no part in the step-size decisions).  Without the flag the solve is
 * controlled by torchdiffeq's DEFAULT mixed norm: max over (rms of the state part, rms of the state-adjoint part, rms of every
 * parameter TENSOR's adjoint part -- each expert scalar, ml_net.weight, ml_net.bias) of the scaled error, the parameter
 * adjoints being those of the whole group (batch-coupled controller, RocheODE up to latent_dim 8; HODE_ERR_UNSUPPORTED else). */
typedef enum hode_flags { HODE_FLAG_HILL2 = 1, HODE_FLAG_ABLATE = 2, HODE_FLAG_ADJ_SEMINORM = 4 } hode_flags;

typedef struct hode_cfg {
    int32_t field;        /* hode_field */
    int32_t latent_dim;   /* D */
    int32_t method;       /* hode_method */
    int32_t controller;   /* hode_controller (dopri5 only) */
    int32_t perturb;      /* fixed-grid option 'perturb' (model.py:825): first/last stage times moved by one ulp */
    int32_t n_dose;       /* doses per patient, columns of dose_t (model.py:507 `times [B, n_dose]`) */
    int32_t expert_grads; /* backward also accumulates the 13 expert-scalar gradients (Roche) */
    int32_t flags;        /* hode_flags */
    double rtol, atol;                /* model.py:1079-1080 */
    double safety, ifactor, dfactor;  /* torchdiffeq defaults 0.9, 10, 0.2 */
    double first_step;                /* <= 0: automatic selection (torchdiffeq `_select_initial_step`) */
    int64_t max_num_steps;            /* torchdiffeq default 2**31-1, counted per output interval */
    int64_t attempt_cap;              /* hard cap on attempts per group for the whole solve (GPU watchdog) */
} hode_cfg;

/* ---- introspection ------------------------------------------------------------------------------------- */
int32_t hode_abi_version(void);
const char* hode_last_error(void);
/* number of floats in one packed parameter set; <0 on unsupported cfg */
int64_t hode_param_count(const hode_cfg* cfg);
/* 1 if (field, latent_dim, method) has a compiled kernel */
int32_t hode_supported(const hode_cfg* cfg);

/* ---- dose schedule: RocheODE.set_action / NeuralODE.set_action (model.py:495-507, 1001-1013) -------------
 * action[t, b] at action + t*stride_t + b*stride_b (element strides, float32).
 * dose_amt[b] = max_t action[t,b];  dose_idx[b, 0..count-1] = indices t with action != 0, ascending, row stride T;
 * dose_count[b] = number of non-zero actions.  (The host checks that all counts are equal, like torch.stack.) */
int32_t hode_dose_schedule(const float* action, int64_t stride_t, int64_t stride_b, int32_t T, int64_t n_traj,
                           float* dose_amt, int32_t* dose_idx, int32_t* dose_count, void* stream);

/* ---- fixed-grid solvers: torchdiffeq FixedGridODESolver.integrate + Euler/Midpoint/RK4 step functions --------
 * (called from model.py:837, 842, 1116 with method in {'euler','midpoint','rk4'}).
 * grid[n_grid]: the solver grid in float32 (host-built exactly like torchdiffeq: `t` itself or
 * arange(niters)*step_size + t[0] with the last point clamped); t_eval[n_t]: output times, float32.
 * dose_t [n_traj, dose_t_stride] float32 dose times (only the first cfg->n_dose columns are read).
 * params [n_param_sets, P]; param_set_of_group[n_groups] (NULL: every group uses set 0).
 * tape (NULL for forward-only): [n_grid-1, n_traj, D] float32, state at the start of every grid step.       */
size_t hode_fixed_tape_bytes(const hode_cfg* cfg, int64_t n_traj, int32_t n_grid);
int32_t hode_fixed_fwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* y0, const float* dose_amt,
                       const float* dose_t, int64_t dose_t_stride, const float* params,
                       const int32_t* param_set_of_group, const float* grid, int32_t n_grid, const float* t_eval,
                       int32_t n_t, float* h_out, float* tape, void* stream);
/* Training forward pass in ONE launch: hode_fixed_fwd with the read-out and the likelihood consumed at every output time
 * -- model.py:1116 (odeint), 1120 (output_function) and 1179 (masked SSE) fused.  While h(t_j) is still in registers the kernel
 * forms x_hat = W h + b, the loss term and ALL gradients of the (scalar) loss that do not need the reverse sweep:
 *   loss [1] = sum (x - x_hat)^2 mask / n_norm,  grad_h [n_t, n_traj, D] = d loss / d h  (the input of hode_fixed_bwd),
 *   grad_w [obs, D], grad_b [obs]  (grad_w may be NULL together with grad_b: forward / evaluation only).
 * h_out may be NULL (the latent solution is then never written), tape as in hode_fixed_fwd.
 * x, mask: [n_t, n_traj, obs] CONTIGUOUS float32, 16-byte aligned (rows travel by TMA bulk copies).  One group, one
 * parameter set.  hode_fixed_fwd_sse_supported() tells whether this (field, D, method, flags, n_dose, obs) has the fused
 * kernel (RocheODE with HODE_FLAG_HILL2, D in {4, 6, 8}, obs in {20, 24, 40, 80}, n_dose == 1); otherwise the call
 * returns HODE_ERR_UNSUPPORTED and the caller issues hode_fixed_fwd + hode_decode_sse. */
int32_t hode_fixed_fwd_sse_supported(const hode_cfg* cfg, int32_t obs, int32_t n_param_sets);
int32_t hode_fixed_fwd_sse(const hode_cfg* cfg, int64_t n_traj, const float* y0, const float* dose_amt, const float* dose_t,
                           int64_t dose_t_stride, const float* params, const float* grid, int32_t n_grid,
                           const float* t_eval, int32_t n_t, const float* W, const float* b, int32_t obs, const float* x,
                           const float* mask, double n_norm, float* h_out, float* tape, float* loss, float* grad_h,
                           float* grad_w, float* grad_b, void* stream);
/* backward (autograd through every step, training_utils.py:50 `loss.backward()`):
 * grad_h [n_t, n_traj, D] -> grad_y0 [n_traj, D] (overwritten), grad_params [n_param_sets, P] (overwritten). */
int32_t hode_fixed_bwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                       const float* dose_t, int64_t dose_t_stride, const float* params,
                       const int32_t* param_set_of_group, int32_t n_param_sets, const float* grid, int32_t n_grid,
                       const float* t_eval, int32_t n_t, const float* grad_h, const float* tape, float* grad_y0,
                       float* grad_params, void* stream);

/* continuous adjoint (torchdiffeq odeint_adjoint -> OdeintAdjointMethod.backward; the import the reference keeps
 * commented out at model.py:9 and north_star's "or the adjoint"): O(1)-memory backward of a fixed-grid solve, no tape.
 * For i = n_t-1 .. 1 the augmented state (y = h[i], a, g_params) is integrated from t[i] back to t[i-1] with the SAME
 * method on torchdiffeq's grid of the negated interval [-t[i], -t[i-1]]; then a += grad_h[i-1], y = h[i-1].
 * h [n_t, n_traj, D]: the forward solution (hode_fixed_fwd's h_out, tape = NULL).
 * adj_grid: the n_t-1 interval grids (negated time, ascending, float32, host-built like `grid` above) back to back,
 * last output interval first; adj_count[iv]: points of interval iv; n_adj_grid = sum(adj_count).
 * grad_y0 / grad_params as in hode_fixed_bwd.  These are the CONTINUOUS-adjoint gradients: they agree with
 * hode_fixed_bwd's discrete gradients to the order of the method, not to rounding. */
int32_t hode_fixed_adjoint(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                           const float* dose_t, int64_t dose_t_stride, const float* params,
                           const int32_t* param_set_of_group, int32_t n_param_sets, const float* adj_grid,
                           int32_t n_adj_grid, const int32_t* adj_count, int32_t n_t, const float* h,
                           const float* grad_h, float* grad_y0, float* grad_params, void* stream);

/* ---- dopri5: torchdiffeq Dopri5Solver (RKAdaptiveStepsizeODESolver.integrate), model.py:1116 ------------------
 * t_eval is float64 (torchdiffeq casts `t` to float64).  controller BATCH: one controller per group
 * (batch <= hode_dopri5_max_batch() = 4096: one CTA up to 512 trajectories, a thread-block cluster of up to 8 CTAs with the
 * group sum exchanged through distributed shared memory beyond); TRAJ: one per trajectory.
 * n_ctrl = n_groups (BATCH) or n_traj (TRAJ).
 * tape (NULL for forward-only): accepted steps, capacity tape_capacity per controller:
 *     tape_t  [n_ctrl, tape_capacity, 2] float64 (t0, dt);  tape_y [tape_capacity, n_traj, D] float32.
 * stats [n_ctrl].                                                                                            */
int64_t hode_dopri5_max_batch(const hode_cfg* cfg);
int32_t hode_dopri5_fwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* y0,
                        const float* dose_amt, const float* dose_t, int64_t dose_t_stride, const float* params,
                        const int32_t* param_set_of_group, const double* t_eval, int32_t n_t, float* h_out,
                        double* tape_t, float* tape_y, int32_t tape_capacity, hode_stats* stats, void* stream);
int32_t hode_dopri5_bwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                        const float* dose_t, int64_t dose_t_stride, const float* params,
                        const int32_t* param_set_of_group, int32_t n_param_sets, const double* t_eval, int32_t n_t,
                        const float* grad_h, const double* tape_t, const float* tape_y, int32_t tape_capacity,
                        const hode_stats* stats, float* grad_y0, float* grad_params, void* stream);

/* Adaptive continuous adjoint: torchdiffeq odeint_adjoint(method='dopri5') -> OdeintAdjointMethod.backward (the import the
 * reference keeps commented out at model.py:9; dopri5 is its default method, sim_config.py:50).  For i = n_t-1 .. 1 the
 * augmented state (y = h[i], a, g_params) is integrated by the dopri5 controller from t[i] back to t[i-1] (negated time),
 * interpolated to t[i-1] by the quartic dense output; then a += grad_h[i-1], y = h[i-1].  No tape.
 * cfg: rtol / atol = the ADJOINT tolerances; controller BATCH (one controller per group, batch <= 512)
 * or TRAJ; flags: HODE_FLAG_ADJ_SEMINORM or none (= mixed norm, see hode_flags).  h [n_t, n_traj, D]: the forward solution (hode_dopri5_fwd without a
 * tape).  stats [n_ctrl]: accepted / rejected attempts summed over the intervals, status as in hode_dopri5_fwd. */
int32_t hode_dopri5_adjoint(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                            const float* dose_t, int64_t dose_t_stride, const float* params,
                            const int32_t* param_set_of_group, int32_t n_param_sets, const double* t_eval, int32_t n_t,
                            const float* h, const float* grad_h, float* grad_y0, float* grad_params, hode_stats* stats,
                            void* stream);

/* ---- decode + masked SSE: output_function (model.py:1097-1100, 1120) and the likelihood of
 * VariationalInference.loss (model.py:1179): loss = sum_{t,b,o} (x - (W h + b))^2 mask / n_norm.
 * x, mask: element strides (st, sb, so).  One pass also produces the unit gradients:
 *   grad_h [n_t, n_traj, D], grad_w [obs, D], grad_b [obs]  (d loss / d .), any of which may be NULL.
 * loss: one float32 (overwritten).  W [obs, D] row-major, b [obs]. */
int32_t hode_decode_sse(int32_t D, int32_t obs, int32_t n_t, int64_t n_traj, double n_norm, const float* h,
                        const float* W, const float* b, const float* x, const float* mask, int64_t st, int64_t sb,
                        int64_t so, float* loss, float* grad_h, float* grad_w, float* grad_b, void* stream);

/* ---- real-data vector fields: RocheODEReal / NeuralODEReal / NeuralODEReal2nd behind DecoderReal.forward
 * (model.py:833-862), which integrates them with the fixed-grid solvers only (experiments/real.sh:9-17: midpoint, rk4;
 * options step_size = 1, perturb = True).  latent_dim in {4, 20} (RocheReal, NeuralReal) / {8, 40} (2nd); hidden <= 64.
 * Packed parameters (float32, hode_real_param_count()):
 *   ROCHE_REAL     : k_immunity, kel, kel2, dx1_net.0.weight [H,3], dx1_net.0.bias [H], dx1_net.2.weight [H],
 *                    dx1_net.2.bias, dx2_net.0.weight [H,2], dx2_net.0.bias [H], dx2_net.2.weight [H], dx2_net.2.bias,
 *                    lin_hh.weight, lin_hz.weight, lin_hr.weight [Z-4, Z-4] each (absent for Z == 4)
 *   NEURAL_REAL(_2ND): ml_net.0.weight [H, Z+1], ml_net.0.bias [H], ml_net.2.weight [OUT, H], ml_net.2.bias [OUT]
 * hode_real_dose_tables replaces set_action_static + the O(T) dose_at_time sums (model.py:646-657, 753-760) by a
 * per-launch table, `tab` [2 if ROCHE_REAL else 1][T+1][n_traj] float32 (ROCHE_REAL reads kel = params[1]); it must be
 * rebuilt whenever the actions or kel change.  One parameter set per call; expert-style scalar gradients always on. */
int64_t hode_real_param_count(int32_t field, int32_t latent_dim, int32_t hidden);
int32_t hode_real_dose_tables(int32_t field, const float* action, int64_t stride_t, int64_t stride_b, int32_t T,
                              int64_t n_traj, const float* params, float* tab, void* stream);
int32_t hode_real_fixed_fwd(int32_t field, int32_t latent_dim, int32_t hidden, int32_t method, int32_t perturb,
                            int64_t n_traj, const float* y0, const float* tab, int32_t T, const float* params,
                            const float* grid, int32_t n_grid, const float* t_eval, int32_t n_t, float* h_out,
                            float* tape, void* stream);
int32_t hode_real_fixed_bwd(int32_t field, int32_t latent_dim, int32_t hidden, int32_t method, int32_t perturb,
                            int64_t n_traj, const float* tab, int32_t T, const float* params, const float* grid,
                            int32_t n_grid, const float* t_eval, int32_t n_t, const float* grad_h, const float* tape,
                            float* grad_y0, float* grad_params, void* stream);

/* ---- Monte-Carlo evaluation (training_utils.py:144-177, evaluate / evaluate_horizon / evaluate_ensemble): the mc_itr
 * decoder solves of a test chunk are ONE solve with n_groups = n_mc (trajectory index s * batch + b); the CRPS that the
 * reference computes with properscoring.crps_ensemble in Python loops is
 *     crps = mean_s |x_s - y| - 1/(2 n_mc^2) sum_{s,s'} |x_s - x_s'|          (n_mc <= 128).
 * hode_crps_ensemble: truth [n], member s of element i at forecasts + i*stride_n + s*stride_mc  ->  out [n].
 * hode_decode_crps:   x_s = W h[t, s*batch + b] + bias evaluated on the fly from the latent solution
 *                     h [n_t, n_mc*batch, D]; observations x with element strides (st, sb, so) -> crps [n_t, batch, obs]. */
int32_t hode_crps_ensemble(const float* truth, const float* forecasts, int64_t n, int32_t n_mc, int64_t stride_n,
                           int64_t stride_mc, float* out, void* stream);
int32_t hode_decode_crps(int32_t D, int32_t obs, int32_t n_t, int64_t batch, int32_t n_mc, const float* h,
                         const float* W, const float* b, const float* x, int64_t st, int64_t sb, int64_t so,
                         float* crps, void* stream);

/* ---- measurement aid: FP32 FMA peak probe (the roofline denominator of the solver kernels; bench.py times it).
 * Launches `blocks` CTAs of 256 threads running 16 independent FFMA chains for `iters` iterations.
 * Returns the number of floating-point operations the launch performs (FMA = 2), or a negative hode_status. */
int64_t hode_bench_ffma(int32_t blocks, int32_t iters, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HODE_H_ */
