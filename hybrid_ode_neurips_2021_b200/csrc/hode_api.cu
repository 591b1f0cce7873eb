// hode_api.cu -- the C ABI of include/hode.h: argument checking and dispatch to the per-(field, D) launchers.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>

#include "hode_bodies.cuh"
#include "hode_real_args.cuh"

namespace hode {
template <class F> int launch_fixed_fwd(const hode_cfg&, const SolveArgs&, cudaStream_t);
template <class F> int launch_fixed_fwd_sse(const hode_cfg&, const SolveArgs&, cudaStream_t);
template <class F> int launch_fixed_bwd(const hode_cfg&, const SolveArgs&, cudaStream_t);
template <class F> int launch_fixed_adj(const hode_cfg&, const SolveArgs&, cudaStream_t);
template <class F> int launch_dopri5_fwd(const hode_cfg&, const SolveArgs&, cudaStream_t);
template <class F> int launch_dopri5_bwd(const hode_cfg&, const SolveArgs&, cudaStream_t);
template <class F> int launch_dopri5_adj(const hode_cfg&, const SolveArgs&, cudaStream_t);
int launch_dose_schedule(const float*, int64_t, int64_t, int32_t, int64_t, float*, int32_t*, int32_t*, cudaStream_t);
int launch_decode_sse(int32_t, int32_t, int32_t, int64_t, double, const float*, const float*, const float*,
                      const float*, const float*, int64_t, int64_t, int64_t, float*, float*, float*, float*,
                      cudaStream_t);
int launch_ffma_probe(int, int, float*, cudaStream_t);
int real_param_count(int, int, int);
int launch_real_dose_table(int, const float*, int64_t, int64_t, int32_t, int64_t, const float*, float*, cudaStream_t);
int launch_real_fixed(bool, int, int, int, const RealArgs&, cudaStream_t);
int launch_crps_ensemble(const float*, const float*, int64_t, int32_t, int64_t, int64_t, float*, cudaStream_t);
int launch_decode_crps(int32_t, int32_t, int32_t, int64_t, int32_t, const float*, const float*, const float*,
                       const float*, int64_t, int64_t, int64_t, float*, cudaStream_t);
}  // namespace hode

using namespace hode;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, const char* a = "", long long b = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

enum Op { OP_FIXED_FWD, OP_FIXED_BWD, OP_DOPRI5_FWD, OP_DOPRI5_BWD, OP_FIXED_ADJ, OP_FIXED_FWD_SSE, OP_DOPRI5_ADJ };

template <class F>
static int run(Op op, const hode_cfg& cfg, const SolveArgs& a, cudaStream_t st) {
    switch (op) {
        case OP_FIXED_FWD: return launch_fixed_fwd<F>(cfg, a, st);
        case OP_FIXED_BWD: return launch_fixed_bwd<F>(cfg, a, st);
        case OP_DOPRI5_FWD: return launch_dopri5_fwd<F>(cfg, a, st);
        case OP_DOPRI5_BWD: return launch_dopri5_bwd<F>(cfg, a, st);
        case OP_FIXED_ADJ: return launch_fixed_adj<F>(cfg, a, st);
        case OP_FIXED_FWD_SSE: return launch_fixed_fwd_sse<F>(cfg, a, st);
        case OP_DOPRI5_ADJ: return launch_dopri5_adj<F>(cfg, a, st);
    }
    return -1;
}

static bool roche_dim_ok(int d) { return d == 4 || d == 6 || d == 8 || d == 12; }

static int dispatch(Op op, const hode_cfg& cfg, const SolveArgs& a, cudaStream_t st) {
    int rc = -1;
    if (cfg.field == HODE_FIELD_ROCHE) {
        const bool hill2 = (cfg.flags & HODE_FLAG_HILL2) != 0;
        const bool ablate = (cfg.flags & HODE_FLAG_ABLATE) != 0;
#define HODE_ROCHE_CASE(DD)                                                                  \
    case DD:                                                                                 \
        rc = ablate ? run<Roche<DD, true, true>>(op, cfg, a, st)                             \
                    : (hill2 ? run<Roche<DD, true>>(op, cfg, a, st) : run<Roche<DD>>(op, cfg, a, st)); \
        break
        switch (cfg.latent_dim) {
            HODE_ROCHE_CASE(4);
            HODE_ROCHE_CASE(6);
            HODE_ROCHE_CASE(8);
            HODE_ROCHE_CASE(12);
            default: return fail(HODE_ERR_UNSUPPORTED, "RocheODE latent_dim %s%lld is not compiled in (4, 6, 8, 12)", "", cfg.latent_dim);
        }
#undef HODE_ROCHE_CASE
    } else if (cfg.field == HODE_FIELD_NEURAL) {
        switch (cfg.latent_dim) {
            case 4: rc = run<Neural<4>>(op, cfg, a, st); break;
            case 6: rc = run<Neural<6>>(op, cfg, a, st); break;
            case 8: rc = run<Neural<8>>(op, cfg, a, st); break;
            case 12: rc = run<Neural<12>>(op, cfg, a, st); break;
            default: return fail(HODE_ERR_UNSUPPORTED, "NeuralODE latent_dim %s%lld is not compiled in (4, 6, 8, 12)", "", cfg.latent_dim);
        }
    } else {
        return fail(HODE_ERR_UNSUPPORTED, "field %s%lld is not compiled in", "", cfg.field);
    }
    if (rc == 0) return HODE_OK;
    if (rc == -2) return fail(HODE_ERR_UNSUPPORTED, "batch-coupled dopri5 group larger than %s%lld trajectories", "",
                              op == OP_DOPRI5_ADJ ? (long long)512 : (long long)hode_dopri5_max_batch(&cfg));
    if (rc == -4) return fail(HODE_ERR_UNSUPPORTED, "the adaptive adjoint with torchdiffeq's default mixed norm is built for the batch-coupled controller and "
                                                    "the RocheODE field up to latent_dim 8; pass adjoint_options={'norm': 'seminorm'} (HODE_FLAG_ADJ_SEMINORM)%s%lld", "", 0);
    if (rc == -3) return fail(HODE_ERR_UNSUPPORTED, "the adaptive adjoint of the NeuralODE field is built for the batch-coupled controller only%s%lld", "", 0);
    if (rc == -1 && op == OP_FIXED_FWD_SSE)
        return fail(HODE_ERR_UNSUPPORTED, "no fused solve + read-out kernel for this field / method / obs / n_dose / parameter-set "
                                          "combination (use hode_fixed_fwd + hode_decode_sse)%s%lld", "", 0);
    if (rc == -1) return fail(HODE_ERR_UNSUPPORTED, "method %s%lld is not valid for this entry point", "", cfg.method);
    if (rc == (int)cudaErrorNoKernelImageForDevice || rc == (int)cudaErrorNoDevice || rc == (int)cudaErrorInsufficientDriver)
        return fail(HODE_ERR_NO_DEVICE, "%s (libhode_b200 holds sm_100a code only)", cudaGetErrorString((cudaError_t)rc));
    return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString((cudaError_t)rc), rc);
}

extern "C" {

int32_t hode_abi_version(void) { return HODE_ABI_VERSION; }
const char* hode_last_error(void) { return g_err; }

int64_t hode_param_count(const hode_cfg* cfg) {
    if (!cfg) return -1;
    const int64_t d = cfg->latent_dim;
    if (cfg->field == HODE_FIELD_ROCHE)
        return d >= 4 ? 13 + (d - 4) * d + (d - 4) + ((cfg->flags & HODE_FLAG_ABLATE) ? 2 : 0) : -1;
    if (cfg->field == HODE_FIELD_NEURAL) return d >= 1 ? 1 + 10 * d * (d + 1) + 10 * d + d * 10 * d + d : -1;
    return -1;
}

int32_t hode_supported(const hode_cfg* cfg) {
    if (!cfg) return 0;
    if (cfg->method < HODE_EULER || cfg->method > HODE_DOPRI5) return 0;
    if (cfg->field == HODE_FIELD_ROCHE || cfg->field == HODE_FIELD_NEURAL) return roche_dim_ok(cfg->latent_dim) ? 1 : 0;
    return 0;
}

int64_t hode_dopri5_max_batch(const hode_cfg* cfg) {
    (void)cfg;
    return 4096;  // forward / reverse sweep: one CTA up to 512 trajectories, a cluster of up to 8 CTAs beyond (hode_launch.cuh)
}

size_t hode_fixed_tape_bytes(const hode_cfg* cfg, int64_t n_traj, int32_t n_grid) {
    if (!cfg || n_traj < 0 || n_grid < 1) return 0;
    return (size_t)(n_grid - 1) * (size_t)n_traj * (size_t)cfg->latent_dim * sizeof(float);
}

static int check_common(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const void* dose_amt, const void* dose_t,
                        int64_t dose_t_stride, const void* params, int32_t n_t) {
    if (!cfg) return fail(HODE_ERR_ARG, "cfg is NULL");
    if (n_groups < 0 || batch < 0) return fail(HODE_ERR_ARG, "negative n_groups/batch");
    if (n_groups * batch > 0 && (!dose_amt || !params)) return fail(HODE_ERR_ARG, "NULL dose_amt/params");
    if (cfg->n_dose < 0 || (cfg->n_dose > 0 && n_groups * batch > 0 && (!dose_t || dose_t_stride < cfg->n_dose)))
        return fail(HODE_ERR_ARG, "bad dose_t / dose_t_stride / n_dose");
    if (n_t < 1) return fail(HODE_ERR_ARG, "n_t must be >= 1");
    if (n_groups > 0x7fffffffLL) return fail(HODE_ERR_ARG, "too many groups");
    return HODE_OK;
}

static void fill_common(SolveArgs& a, const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                        const float* dose_t, int64_t dose_t_stride, const float* params, const int32_t* pset) {
    memset(&a, 0, sizeof(a));
    a.n_groups = n_groups; a.batch = batch;
    a.dose_amt = dose_amt; a.dose_t = dose_t; a.dose_t_stride = dose_t_stride; a.n_dose = cfg->n_dose;
    a.params = params; a.pset = pset; a.perturb = cfg->perturb;
    a.rtol_f = (float)cfg->rtol; a.atol_f = (float)cfg->atol;
    a.safety = cfg->safety; a.ifactor = cfg->ifactor; a.dfactor = cfg->dfactor; a.first_step = cfg->first_step;
    a.max_num_steps = cfg->max_num_steps; a.attempt_cap = cfg->attempt_cap;
    a.per_traj = cfg->controller == HODE_CTRL_TRAJ;
}

int32_t hode_dose_schedule(const float* action, int64_t stride_t, int64_t stride_b, int32_t T, int64_t n_traj,
                           float* dose_amt, int32_t* dose_idx, int32_t* dose_count, void* stream) {
    if (T < 1 || n_traj < 0) return fail(HODE_ERR_ARG, "bad T / n_traj");
    if (n_traj == 0) return HODE_OK;
    if (!action || !dose_amt || !dose_idx || !dose_count) return fail(HODE_ERR_ARG, "NULL pointer");
    const int rc = launch_dose_schedule(action, stride_t, stride_b, T, n_traj, dose_amt, dose_idx, dose_count,
                                        (cudaStream_t)stream);
    if (rc != 0) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString((cudaError_t)rc), rc);
    return HODE_OK;
}

int32_t hode_fixed_fwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* y0, const float* dose_amt,
                       const float* dose_t, int64_t dose_t_stride, const float* params,
                       const int32_t* param_set_of_group, const float* grid, int32_t n_grid, const float* t_eval,
                       int32_t n_t, float* h_out, float* tape, void* stream) {
    int rc = check_common(cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, n_t);
    if (rc) return rc;
    if (cfg->method == HODE_DOPRI5) return fail(HODE_ERR_ARG, "hode_fixed_fwd called with method dopri5");
    if (n_grid < 1 || !grid || !t_eval) return fail(HODE_ERR_ARG, "bad grid / t_eval");
    if (n_groups * batch == 0) return HODE_OK;
    if (!y0 || !h_out) return fail(HODE_ERR_ARG, "NULL y0 / h_out");
    SolveArgs a;
    fill_common(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, param_set_of_group);
    a.y0 = y0; a.grid = grid; a.n_grid = n_grid; a.t_eval_f = t_eval; a.n_t = n_t; a.h_out = h_out; a.tape_y = tape;
    return dispatch(OP_FIXED_FWD, *cfg, a, (cudaStream_t)stream);
}

int32_t hode_fixed_fwd_sse_supported(const hode_cfg* cfg, int32_t obs, int32_t n_param_sets) {
    if (!cfg || cfg->field != HODE_FIELD_ROCHE || cfg->method == HODE_DOPRI5) return 0;
    if (!(cfg->flags & HODE_FLAG_HILL2) || (cfg->flags & HODE_FLAG_ABLATE)) return 0;
    const int d = cfg->latent_dim;
    if (d != 4 && d != 6 && d != 8) return 0;
    if (cfg->n_dose != 1 || n_param_sets != 1) return 0;
    return (obs == 20 || obs == 24 || obs == 40 || obs == 80) ? 1 : 0;  // the reference's observation widths
}

int32_t hode_fixed_fwd_sse(const hode_cfg* cfg, int64_t n_traj, const float* y0, const float* dose_amt, const float* dose_t,
                           int64_t dose_t_stride, const float* params, const float* grid, int32_t n_grid,
                           const float* t_eval, int32_t n_t, const float* W, const float* b, int32_t obs, const float* x,
                           const float* mask, double n_norm, float* h_out, float* tape, float* loss, float* grad_h,
                           float* grad_w, float* grad_b, void* stream) {
    int rc = check_common(cfg, 1, n_traj, dose_amt, dose_t, dose_t_stride, params, n_t);
    if (rc) return rc;
    if (!hode_fixed_fwd_sse_supported(cfg, obs, 1))
        return fail(HODE_ERR_UNSUPPORTED, "no fused solve + read-out kernel for this configuration (use hode_fixed_fwd + hode_decode_sse)");
    if (n_grid < 1 || !grid || !t_eval || !W || !b || !loss) return fail(HODE_ERR_ARG, "bad grid / t_eval / W / b / loss");
    if (((uintptr_t)x | (uintptr_t)mask) & 15) return fail(HODE_ERR_ARG, "x / mask must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
    if (e == cudaSuccess && grad_w) e = cudaMemsetAsync(grad_w, 0, sizeof(float) * (size_t)obs * cfg->latent_dim, st);
    if (e == cudaSuccess && grad_b) e = cudaMemsetAsync(grad_b, 0, sizeof(float) * (size_t)obs, st);
    if (e != cudaSuccess) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString(e), (long long)e);
    if (n_traj == 0) return HODE_OK;
    if (!y0 || !x || !mask) return fail(HODE_ERR_ARG, "NULL y0 / x / mask");
    if (grad_b && !grad_w) return fail(HODE_ERR_ARG, "grad_b without grad_w");
    SolveArgs a;
    fill_common(a, cfg, 1, n_traj, dose_amt, dose_t, dose_t_stride, params, nullptr);
    a.y0 = y0; a.grid = grid; a.n_grid = n_grid; a.t_eval_f = t_eval; a.n_t = n_t; a.h_out = h_out; a.tape_y = tape;
    a.sse_x = x; a.sse_mask = mask; a.sse_w = W; a.sse_b = b; a.sse_obs = obs;
    a.sse_scale = (float)(-2.0 / n_norm); a.sse_inv_norm = (float)(1.0 / n_norm);
    a.sse_loss = loss; a.sse_grad_h = grad_h; a.sse_grad_w = grad_w; a.sse_grad_b = grad_b;
    return dispatch(OP_FIXED_FWD_SSE, *cfg, a, st);
}

int32_t hode_fixed_bwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                       const float* dose_t, int64_t dose_t_stride, const float* params,
                       const int32_t* param_set_of_group, int32_t n_param_sets, const float* grid, int32_t n_grid,
                       const float* t_eval, int32_t n_t, const float* grad_h, const float* tape, float* grad_y0,
                       float* grad_params, void* stream) {
    int rc = check_common(cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, n_t);
    if (rc) return rc;
    if (cfg->method == HODE_DOPRI5) return fail(HODE_ERR_ARG, "hode_fixed_bwd called with method dopri5");
    if (n_grid < 1 || !grid || !t_eval || n_param_sets < 1 || !grad_params) return fail(HODE_ERR_ARG, "bad grid / t_eval / grad_params");
    const int64_t P = hode_param_count(cfg);
    cudaError_t e = cudaMemsetAsync(grad_params, 0, sizeof(float) * (size_t)P * (size_t)n_param_sets, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString(e), (long long)e);
    if (n_groups * batch == 0) return HODE_OK;
    if (!grad_h || !grad_y0 || (n_grid > 1 && !tape)) return fail(HODE_ERR_ARG, "NULL grad_h / grad_y0 / tape");
    SolveArgs a;
    fill_common(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, param_set_of_group);
    a.n_param_sets = n_param_sets;
    a.grid = grid; a.n_grid = n_grid; a.t_eval_f = t_eval; a.n_t = n_t;
    a.grad_h = grad_h; a.tape_y = const_cast<float*>(tape); a.grad_y0 = grad_y0; a.grad_params = grad_params;
    return dispatch(OP_FIXED_BWD, *cfg, a, (cudaStream_t)stream);
}

int32_t hode_fixed_adjoint(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                           const float* dose_t, int64_t dose_t_stride, const float* params,
                           const int32_t* param_set_of_group, int32_t n_param_sets, const float* adj_grid,
                           int32_t n_adj_grid, const int32_t* adj_count, int32_t n_t, const float* h,
                           const float* grad_h, float* grad_y0, float* grad_params, void* stream) {
    int rc = check_common(cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, n_t);
    if (rc) return rc;
    if (cfg->method == HODE_DOPRI5) return fail(HODE_ERR_UNSUPPORTED, "the continuous adjoint is built for the fixed-grid methods only (method %s%lld)", "", cfg->method);
    if (n_adj_grid < 0 || n_param_sets < 1 || !grad_params || (n_t > 1 && (!adj_grid || !adj_count)))
        return fail(HODE_ERR_ARG, "bad adj_grid / adj_count / grad_params");
    const int64_t P = hode_param_count(cfg);
    cudaError_t e = cudaMemsetAsync(grad_params, 0, sizeof(float) * (size_t)P * (size_t)n_param_sets, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString(e), (long long)e);
    if (n_groups * batch == 0) return HODE_OK;
    if (!h || !grad_h || !grad_y0) return fail(HODE_ERR_ARG, "NULL h / grad_h / grad_y0");
    SolveArgs a;
    fill_common(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, param_set_of_group);
    a.n_param_sets = n_param_sets;
    a.grid = adj_grid; a.n_grid = n_adj_grid; a.adj_cnt = adj_count; a.n_t = n_t;
    a.h_out = const_cast<float*>(h); a.grad_h = grad_h; a.grad_y0 = grad_y0; a.grad_params = grad_params;
    return dispatch(OP_FIXED_ADJ, *cfg, a, (cudaStream_t)stream);
}

int32_t hode_dopri5_adjoint(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                            const float* dose_t, int64_t dose_t_stride, const float* params,
                            const int32_t* param_set_of_group, int32_t n_param_sets, const double* t_eval, int32_t n_t,
                            const float* h, const float* grad_h, float* grad_y0, float* grad_params, hode_stats* stats,
                            void* stream) {
    int rc = check_common(cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, n_t);
    if (rc) return rc;
    if (cfg->method != HODE_DOPRI5) return fail(HODE_ERR_ARG, "hode_dopri5_adjoint called with a fixed-grid method (use hode_fixed_adjoint)");
    if (!t_eval || !stats || n_param_sets < 1 || !grad_params) return fail(HODE_ERR_ARG, "NULL / bad t_eval, stats or grad_params");
    const int64_t P = hode_param_count(cfg);
    cudaError_t e = cudaMemsetAsync(grad_params, 0, sizeof(float) * (size_t)P * (size_t)n_param_sets, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString(e), (long long)e);
    if (n_groups * batch == 0) return HODE_OK;
    if (!h || !grad_h || !grad_y0) return fail(HODE_ERR_ARG, "NULL h / grad_h / grad_y0");
    SolveArgs a;
    fill_common(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, param_set_of_group);
    a.n_param_sets = n_param_sets;
    a.t_eval_d = t_eval; a.n_t = n_t; a.h_out = const_cast<float*>(h); a.grad_h = grad_h;
    a.stats = stats; a.grad_y0 = grad_y0; a.grad_params = grad_params;
    a.adj_mixed = (cfg->flags & HODE_FLAG_ADJ_SEMINORM) ? 0 : 1;  // torchdiffeq's default is the mixed norm
    return dispatch(OP_DOPRI5_ADJ, *cfg, a, (cudaStream_t)stream);
}

int32_t hode_dopri5_fwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* y0,
                        const float* dose_amt, const float* dose_t, int64_t dose_t_stride, const float* params,
                        const int32_t* param_set_of_group, const double* t_eval, int32_t n_t, float* h_out,
                        double* tape_t, float* tape_y, int32_t tape_capacity, hode_stats* stats, void* stream) {
    int rc = check_common(cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, n_t);
    if (rc) return rc;
    if (cfg->method != HODE_DOPRI5) return fail(HODE_ERR_ARG, "hode_dopri5_fwd called with a fixed-grid method");
    if (!t_eval || !stats) return fail(HODE_ERR_ARG, "NULL t_eval / stats");
    if ((tape_y == nullptr) != (tape_t == nullptr)) return fail(HODE_ERR_ARG, "tape_t and tape_y must both be given or both be NULL");
    if (tape_y && tape_capacity < 1) return fail(HODE_ERR_ARG, "tape_capacity must be >= 1");
    if (n_groups * batch == 0) return HODE_OK;
    if (!y0 || !h_out) return fail(HODE_ERR_ARG, "NULL y0 / h_out");
    SolveArgs a;
    fill_common(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, param_set_of_group);
    a.y0 = y0; a.t_eval_d = t_eval; a.n_t = n_t; a.h_out = h_out;
    a.tape_t = tape_t; a.tape_y = tape_y; a.tape_cap = tape_capacity; a.stats = stats;
    return dispatch(OP_DOPRI5_FWD, *cfg, a, (cudaStream_t)stream);
}

int32_t hode_dopri5_bwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                        const float* dose_t, int64_t dose_t_stride, const float* params,
                        const int32_t* param_set_of_group, int32_t n_param_sets, const double* t_eval, int32_t n_t,
                        const float* grad_h, const double* tape_t, const float* tape_y, int32_t tape_capacity,
                        const hode_stats* stats, float* grad_y0, float* grad_params, void* stream) {
    int rc = check_common(cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, n_t);
    if (rc) return rc;
    if (cfg->method != HODE_DOPRI5) return fail(HODE_ERR_ARG, "hode_dopri5_bwd called with a fixed-grid method");
    if (!t_eval || !stats || !tape_t || !tape_y || n_param_sets < 1 || !grad_params || tape_capacity < 1)
        return fail(HODE_ERR_ARG, "NULL / bad t_eval, stats, tape or grad_params");
    const int64_t P = hode_param_count(cfg);
    cudaError_t e = cudaMemsetAsync(grad_params, 0, sizeof(float) * (size_t)P * (size_t)n_param_sets, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString(e), (long long)e);
    if (n_groups * batch == 0) return HODE_OK;
    if (!grad_h || !grad_y0) return fail(HODE_ERR_ARG, "NULL grad_h / grad_y0");
    SolveArgs a;
    fill_common(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, param_set_of_group);
    a.n_param_sets = n_param_sets;
    a.t_eval_d = t_eval; a.n_t = n_t; a.grad_h = grad_h;
    a.tape_t = const_cast<double*>(tape_t); a.tape_y = const_cast<float*>(tape_y); a.tape_cap = tape_capacity;
    a.stats = const_cast<hode_stats*>(stats); a.grad_y0 = grad_y0; a.grad_params = grad_params;
    return dispatch(OP_DOPRI5_BWD, *cfg, a, (cudaStream_t)stream);
}

int32_t hode_decode_sse(int32_t D, int32_t obs, int32_t n_t, int64_t n_traj, double n_norm, const float* h,
                        const float* W, const float* b, const float* x, const float* mask, int64_t st, int64_t sb,
                        int64_t so, float* loss, float* grad_h, float* grad_w, float* grad_b, void* stream) {
    if (D < 1 || D > 64 || obs < 1 || n_t < 1 || n_traj < 0) return fail(HODE_ERR_ARG, "bad D / obs / n_t / n_traj");
    if (!loss || !W || !b) return fail(HODE_ERR_ARG, "NULL loss / W / b");
    if (n_traj > 0 && (!h || !x || !mask)) return fail(HODE_ERR_ARG, "NULL h / x / mask");
    const int rc = launch_decode_sse(D, obs, n_t, n_traj, n_norm, h, W, b, x, mask, st, sb, so, loss, grad_h, grad_w,
                                     grad_b, (cudaStream_t)stream);
    if (rc == -1) return fail(HODE_ERR_UNSUPPORTED, "decode_sse: obs*D too large for one CTA's shared memory");
    if (rc != 0) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString((cudaError_t)rc), rc);
    return HODE_OK;
}

int64_t hode_real_param_count(int32_t field, int32_t latent_dim, int32_t hidden) {
    return real_param_count(field, latent_dim, hidden);
}

int32_t hode_real_dose_tables(int32_t field, const float* action, int64_t stride_t, int64_t stride_b, int32_t T,
                              int64_t n_traj, const float* params, float* tab, void* stream) {
    if (field < HODE_FIELD_ROCHE_REAL || field > HODE_FIELD_NEURAL_REAL_2ND) return fail(HODE_ERR_ARG, "not a real-data field");
    if (T < 1 || n_traj < 0) return fail(HODE_ERR_ARG, "bad T / n_traj");
    if (n_traj == 0) return HODE_OK;
    if (!action || !tab || (field == HODE_FIELD_ROCHE_REAL && !params)) return fail(HODE_ERR_ARG, "NULL pointer");
    const int rc = launch_real_dose_table(field, action, stride_t, stride_b, T, n_traj, params, tab, (cudaStream_t)stream);
    if (rc != 0) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString((cudaError_t)rc), rc);
    return HODE_OK;
}

static int real_common(bool bwd, int32_t field, int32_t latent_dim, int32_t hidden, int32_t method, int32_t perturb,
                       int64_t n_traj, const float* tab, int32_t T, const float* params, const float* grid, int32_t n_grid,
                       const float* t_eval, int32_t n_t, RealArgs& r) {
    const int P = real_param_count(field, latent_dim, hidden);
    if (P < 0) return fail(HODE_ERR_UNSUPPORTED, "real-data field / latent_dim / hidden (<= 64) %s%lld is not compiled in", "", latent_dim);
    if (method < HODE_EULER || method > HODE_RK4_38) return fail(HODE_ERR_UNSUPPORTED, "real-data fields are integrated with euler / midpoint / rk4 only (method %s%lld)", "", method);
    if (n_traj < 0 || n_grid < 1 || n_t < 1 || T < 1) return fail(HODE_ERR_ARG, "bad sizes");
    if (!tab || !params || !grid || !t_eval) return fail(HODE_ERR_ARG, "NULL pointer");
    memset(&r, 0, sizeof(r));
    r.a.n_groups = 1; r.a.batch = n_traj; r.a.params = params; r.a.perturb = perturb;
    r.a.grid = grid; r.a.n_grid = n_grid; r.a.t_eval_f = t_eval; r.a.n_t = n_t;
    r.tab = tab; r.T = T; r.hidden = hidden; r.P = P;
    (void)bwd;
    return HODE_OK;
}

int32_t hode_real_fixed_fwd(int32_t field, int32_t latent_dim, int32_t hidden, int32_t method, int32_t perturb,
                            int64_t n_traj, const float* y0, const float* tab, int32_t T, const float* params,
                            const float* grid, int32_t n_grid, const float* t_eval, int32_t n_t, float* h_out,
                            float* tape, void* stream) {
    RealArgs r;
    int rc = real_common(false, field, latent_dim, hidden, method, perturb, n_traj, tab, T, params, grid, n_grid, t_eval, n_t, r);
    if (rc) return rc;
    if (n_traj == 0) return HODE_OK;
    if (!y0 || !h_out) return fail(HODE_ERR_ARG, "NULL y0 / h_out");
    r.a.y0 = y0; r.a.h_out = h_out; r.a.tape_y = tape;
    rc = launch_real_fixed(false, field, latent_dim, method, r, (cudaStream_t)stream);
    if (rc != 0) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString((cudaError_t)rc), rc);
    return HODE_OK;
}

int32_t hode_real_fixed_bwd(int32_t field, int32_t latent_dim, int32_t hidden, int32_t method, int32_t perturb,
                            int64_t n_traj, const float* tab, int32_t T, const float* params, const float* grid,
                            int32_t n_grid, const float* t_eval, int32_t n_t, const float* grad_h, const float* tape,
                            float* grad_y0, float* grad_params, void* stream) {
    RealArgs r;
    int rc = real_common(true, field, latent_dim, hidden, method, perturb, n_traj, tab, T, params, grid, n_grid, t_eval, n_t, r);
    if (rc) return rc;
    if (!grad_params) return fail(HODE_ERR_ARG, "NULL grad_params");
    cudaError_t e = cudaMemsetAsync(grad_params, 0, sizeof(float) * (size_t)r.P, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString(e), (long long)e);
    if (n_traj == 0) return HODE_OK;
    if (!grad_h || !grad_y0 || (n_grid > 1 && !tape)) return fail(HODE_ERR_ARG, "NULL grad_h / grad_y0 / tape");
    r.a.grad_h = grad_h; r.a.tape_y = const_cast<float*>(tape); r.a.grad_y0 = grad_y0; r.a.grad_params = grad_params;
    rc = launch_real_fixed(true, field, latent_dim, method, r, (cudaStream_t)stream);
    if (rc != 0) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString((cudaError_t)rc), rc);
    return HODE_OK;
}

int32_t hode_crps_ensemble(const float* truth, const float* forecasts, int64_t n, int32_t n_mc, int64_t stride_n,
                           int64_t stride_mc, float* out, void* stream) {
    if (n < 0 || n_mc < 1) return fail(HODE_ERR_ARG, "bad n / n_mc");
    if (n == 0) return HODE_OK;
    if (!truth || !forecasts || !out) return fail(HODE_ERR_ARG, "NULL pointer");
    const int rc = launch_crps_ensemble(truth, forecasts, n, n_mc, stride_n, stride_mc, out, (cudaStream_t)stream);
    if (rc == -1) return fail(HODE_ERR_UNSUPPORTED, "crps_ensemble: more than %s%lld ensemble members", "", 128);
    if (rc != 0) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString((cudaError_t)rc), rc);
    return HODE_OK;
}

int32_t hode_decode_crps(int32_t D, int32_t obs, int32_t n_t, int64_t batch, int32_t n_mc, const float* h,
                         const float* W, const float* b, const float* x, int64_t st, int64_t sb, int64_t so,
                         float* crps, void* stream) {
    if (D < 1 || obs < 1 || n_t < 1 || batch < 0 || n_mc < 1) return fail(HODE_ERR_ARG, "bad D / obs / n_t / batch / n_mc");
    if (batch == 0) return HODE_OK;
    if (!h || !W || !b || !x || !crps) return fail(HODE_ERR_ARG, "NULL pointer");
    const int rc = launch_decode_crps(D, obs, n_t, batch, n_mc, h, W, b, x, st, sb, so, crps, (cudaStream_t)stream);
    if (rc == -1) return fail(HODE_ERR_UNSUPPORTED, "decode_crps: latent_dim / n_mc (<= 128) / size not compiled in");
    if (rc != 0) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString((cudaError_t)rc), rc);
    return HODE_OK;
}

int64_t hode_bench_ffma(int32_t blocks, int32_t iters, float* out, void* stream) {
    if (blocks < 1 || iters < 1 || !out) return fail(HODE_ERR_ARG, "bad blocks / iters / out");
    const int rc = launch_ffma_probe(blocks, iters, out, (cudaStream_t)stream);
    if (rc != 0) return fail(HODE_ERR_CUDA, "CUDA error: %s (%lld)", cudaGetErrorString((cudaError_t)rc), rc);
    return (int64_t)2 * 16 * (int64_t)iters * 256 * (int64_t)blocks;
}

}  // extern "C"
