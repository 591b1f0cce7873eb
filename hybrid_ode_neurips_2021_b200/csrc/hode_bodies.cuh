// hode_bodies.cuh -- whole-solve bodies executed by one thread per trajectory: fixed-grid forward / reverse sweep,
// dopri5 forward (batch-coupled or per-trajectory controller) / reverse sweep over the tape of accepted steps.
// `Comm` supplies the only cross-thread operation the solvers need: a sum over the controller group.
#pragma once
#include "hode_core.cuh"
#include "../../include/hode.h"

namespace hode {

struct SolveArgs {
    int64_t n_groups, batch;
    const float* y0;
    const float* dose_amt;
    const float* dose_t;
    int64_t dose_t_stride;
    int32_t n_dose;
    const float* params;
    const int32_t* pset;
    int32_t n_param_sets;
    int32_t perturb;
    // fixed grid
    const float* grid;
    int32_t n_grid;
    const float* t_eval_f;
    // dopri5
    const double* t_eval_d;
    double* tape_t;
    int32_t tape_cap;
    hode_stats* stats;
    float rtol_f, atol_f;
    double safety, ifactor, dfactor, first_step;
    int64_t max_num_steps, attempt_cap;
    int32_t per_traj;
    int64_t ctrl_batch;  // > 0: "flat" launch (threads enumerate all trajectories); trajectories per controller group
    // common
    int32_t n_t;
    float* h_out;
    float* tape_y;
    // backward
    const float* grad_h;
    float* grad_y0;
    float* grad_params;
    // fused read-out + masked SSE (hode_fixed_fwd_sse): x / mask [n_t, n_traj, obs] contiguous float32, W [obs, D], b [obs]
    const float* sse_x;
    const float* sse_mask;
    const float* sse_w;
    const float* sse_b;
    int32_t sse_obs;
    float sse_scale, sse_inv_norm;  // -2 / n_norm, 1 / n_norm
    float* sse_loss;
    float* sse_grad_h;
    float* sse_grad_w;
    float* sse_grad_b;
    int adj_mixed = 0;  // adaptive adjoint: 1 = torchdiffeq's default mixed norm (the parameter adjoints take part in the error control)
    // continuous adjoint: `grid` holds the n_t - 1 reversed-time interval grids back to back (n_grid points in total),
    // adj_cnt[iv] = number of grid points of interval iv (iv = 0 is the LAST output interval)
    const int32_t* adj_cnt;
};

// Reverse sweeps request the tape entry of the NEXT step to be reversed one step ahead (D more live registers).  Only where
// registers allow: RocheODE with <= 64 packed parameters (D <= 8).  At D = 12 (117 accumulators, kernels already spill) the
// extra live state costs more than the hidden latency (measured: dopri5 reverse sweep 11.9 -> 13.4 ms at the C3 shape).
template <class F> struct TapePrefetch { static constexpr bool value = false; };
template <int D_, bool H_, bool A_> struct TapePrefetch<Roche<D_, H_, A_>> { static constexpr bool value = Roche<D_, H_, A_>::P <= 64; };
// dopri5 reverse sweep: always for the RocheODE field (at D = 12 the stage rows live in shared memory and the cooperative
// accumulators take 14 registers, so the D + 4 registers of the prefetched entry fit)
template <class F> struct TapePrefetchD5 { static constexpr bool value = false; };
template <int D_, bool H_, bool A_> struct TapePrefetchD5<Roche<D_, H_, A_>> { static constexpr bool value = true; };

// ---- vector load/store of one trajectory's D contiguous floats ------------------------------------------------
template <int D>
HODE_HD void load_vec(const float* __restrict__ p, float (&v)[D]) {
#if HODE_DEVICE_BUILD
    if (D % 4 == 0) {
#pragma unroll
        for (int i = 0; i < D / 4; ++i) {
            const float4 q = reinterpret_cast<const float4*>(p)[i];
            v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
        }
        return;
    } else if (D % 2 == 0) {
#pragma unroll
        for (int i = 0; i < D / 2; ++i) {
            const float2 q = reinterpret_cast<const float2*>(p)[i];
            v[2 * i] = q.x; v[2 * i + 1] = q.y;
        }
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = p[i];
}
template <int D>
HODE_HD void store_vec(float* __restrict__ p, const float (&v)[D]) {
#if HODE_DEVICE_BUILD
    if (D % 4 == 0) {
#pragma unroll
        for (int i = 0; i < D / 4; ++i)
            reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        return;
    } else if (D % 2 == 0) {
#pragma unroll
        for (int i = 0; i < D / 2; ++i) reinterpret_cast<float2*>(p)[i] = make_float2(v[2 * i], v[2 * i + 1]);
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < D; ++i) p[i] = v[i];
}

// ==============================================================================================================
// fixed grid (tde solvers.py FixedGridODESolver.integrate)
// ==============================================================================================================
// Where the solution at an output time goes.  HSink: into h_out [n_t, n_traj, D] (what torchdiffeq.odeint returns).  The
// training kernels use a sink that consumes h(t_j) on the spot -- read-out, masked squared error and its gradients
// (SseSink, hode_sse.cuh) -- so that the latent solution is not written and re-read at all.
template <int D>
struct HSink {
    float* h_out;
    int64_t n_traj, idx;
    HODE_HD void emit(int j, const float (&v)[D]) { store_vec<D>(h_out + ((int64_t)j * n_traj + idx) * D, v); }
};

template <class F, int METHOD, class PS, class Dose, class Sink>
HODE_HD void fixed_fwd_traj(const SolveArgs& a, PS sp, const Dose& ds, int64_t idx, Sink& sink, bool write_tape = true) {
    constexpr int D = F::D;
    const int64_t n_traj = a.n_groups * a.batch;
    float y[D], y1[D];
    load_vec<D>(a.y0 + idx * D, y);
#pragma unroll
    for (int d = 0; d < D; ++d) y1[d] = y[d];
    const bool perturb = a.perturb != 0;
    // loop-carried scalars instead of per-step index arithmetic: grid time, tape cursor, next output time
    float t0 = a.grid[0], t1 = t0;
    float* tp = (a.tape_y != nullptr && write_tape) ? a.tape_y + idx * D : nullptr;
    const int64_t tape_stride = n_traj * D;
    int j = 0, s = 0;
    float tj = a.t_eval_f[0];  // == grid[0]: solution[0] = y0 is emitted by the first trip below
    // The grid is walked in two nested loops.  The inner one takes steps until an output time is reached and contains nothing
    // but the step; the outer one holds the ONE place where outputs are emitted (y at t0, y1 at t1, or the linear
    // interpolant).  A sink that does real work (SseSink) therefore appears once in the code and outside the hot loop.
    for (;;) {
        if (t1 >= tj) {  // false for NaN times, like the reference's `while`
            while (j < a.n_t && t1 >= a.t_eval_f[j]) {
                const float te = a.t_eval_f[j];
                float v[D];
                if (te == t0) {
#pragma unroll
                    for (int d = 0; d < D; ++d) v[d] = y[d];
                } else if (te == t1) {
#pragma unroll
                    for (int d = 0; d < D; ++d) v[d] = y1[d];
                } else {  // _linear_interp
                    const float slope = div_rn(sub_rn(te, t0), sub_rn(t1, t0));
#pragma unroll
                    for (int d = 0; d < D; ++d) v[d] = y[d] + slope * (y1[d] - y[d]);
                }
                sink.emit(j, v);
                ++j;
            }
            tj = (j < a.n_t) ? a.t_eval_f[j] : INFINITY;
            if (s == 0 && !F::params_ok(sp)) {  // kernel variant and parameters disagree (hode_cfg.flags): fail loudly
#pragma unroll
                for (int d = 0; d < D; ++d) y1[d] = nanf("");
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) y[d] = y1[d];
        t0 = t1;
        if (s + 1 >= a.n_grid) break;
        for (;;) {
            t1 = a.grid[s + 1];
            const float dt = sub_rn(t1, t0);
            if (tp != nullptr) {
                store_vec<D>(tp, y);
                tp += tape_stride;
            }
            fixed_step<F, METHOD>(sp, ds, t0, t1, dt, perturb, y, y1);
            ++s;
            if (t1 >= tj || s + 1 >= a.n_grid) break;
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = y1[d];
            t0 = t1;
        }
    }
}
// default sink: the latent solution
template <class F, int METHOD, class PS, class Dose>
HODE_HD void fixed_fwd_traj(const SolveArgs& a, PS sp, const Dose& ds, int64_t idx) {
    HSink<F::D> sink{a.h_out, a.n_groups * a.batch, idx};
    fixed_fwd_traj<F, METHOD>(a, sp, ds, idx, sink);
}

template <class F, int METHOD, bool EG, class PS, class Dose, class ACC>
HODE_HD void fixed_bwd_traj(const SolveArgs& a, PS sp, const Dose& ds, int64_t idx, ACC acc, bool valid = true) {
    constexpr int D = F::D;
    const int64_t n_traj = a.n_groups * a.batch;
    float lam[D];
#pragma unroll
    for (int d = 0; d < D; ++d) lam[d] = 0.0f;
    int j = a.n_t - 1;
    const bool perturb = a.perturb != 0;
    // loop-carried scalars: grid time, tape cursor, time of the latest output not yet consumed
    const int64_t tape_stride = n_traj * D;
    const float* tp = a.tape_y + ((int64_t)(a.n_grid - 2) * n_traj + idx) * D;
    float t1 = a.n_grid >= 1 ? a.grid[a.n_grid - 1] : 0.0f;
    float tj = (j >= 1) ? a.t_eval_f[j] : -INFINITY;
    // The tape entry of step s-1 is requested while step s is being reversed: with 3 warps per scheduler nothing else
    // hides an HBM round trip per step (ncu: stall_long_scoreboard 1.3 per issue without this, 0.2 in the tape-free
    // adjoint kernel).
    constexpr bool PF = TapePrefetch<F>::value;
    float ynext[D];
    if (PF) {
        if (a.n_grid >= 2) load_vec<D>(tp, ynext);
        tp -= tape_stride;
    }
    for (int s = a.n_grid - 2; s >= 0; --s) {
        const float t0 = a.grid[s];
        const float dt = sub_rn(t1, t0);
        float y0[D], yb0[D], lam0[D];
        if (PF) {
#pragma unroll
            for (int d = 0; d < D; ++d) y0[d] = ynext[d];
            if (s > 0) load_vec<D>(tp, ynext);
        } else {
            load_vec<D>(tp, y0);
        }
        tp -= tape_stride;
#pragma unroll
        for (int d = 0; d < D; ++d) yb0[d] = 0.0f;
        // outputs emitted by this step in the forward pass: grid[s] < t_eval[j] <= grid[s+1]
        if (tj > t0) {
            while (j >= 1 && a.t_eval_f[j] > t0) {
                const float te = a.t_eval_f[j];
                float g[D];
                load_vec<D>(a.grad_h + ((int64_t)j * n_traj + idx) * D, g);
                if (!valid) {
#pragma unroll
                    for (int d = 0; d < D; ++d) g[d] = 0.0f;
                }
                if (te == t1) {
#pragma unroll
                    for (int d = 0; d < D; ++d) lam[d] += g[d];
                } else {
                    const float slope = div_rn(sub_rn(te, t0), sub_rn(t1, t0));
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        lam[d] += slope * g[d];
                        yb0[d] += g[d] - slope * g[d];
                    }
                }
                --j;
            }
            tj = (j >= 1) ? a.t_eval_f[j] : -INFINITY;
        }
        fixed_step_vjp<F, METHOD, EG>(sp, ds, t0, t1, dt, perturb, y0, lam, lam0, acc);
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] = lam0[d] + yb0[d];
        t1 = t0;
    }
    if (!valid) return;
    float g0[D];
    load_vec<D>(a.grad_h + idx * D, g0);
#pragma unroll
    for (int d = 0; d < D; ++d) lam[d] += g0[d];
    if (!F::params_ok(sp)) {
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] = nanf("");
    }
    store_vec<D>(a.grad_y0 + idx * D, lam);
}

// ==============================================================================================================
// continuous adjoint of a fixed-grid solve (tde adjoint.py OdeintAdjointMethod.backward): for i = n_t-1 .. 1 the
// augmented state (y = h[i], a, g_theta) is integrated from t[i] back to t[i-1] on the solver grid of the NEGATED
// interval; then a += grad_h[i-1] and y is reset to the forward solution h[i-1].  O(1) memory: no tape.
// ==============================================================================================================
template <class F, int METHOD, bool EG, class PS, class Dose, class ACC>
HODE_HD void fixed_adj_traj(const SolveArgs& a, PS sp, const Dose& ds, int64_t idx, ACC acc, bool valid = true) {
    constexpr int D = F::D;
    const int64_t n_traj = a.n_groups * a.batch;
    const bool perturb = a.perturb != 0;
    float lam[D], y[D];
#pragma unroll
    for (int d = 0; d < D; ++d) lam[d] = 0.0f;
    const float* gp = a.grid;
    for (int iv = 0; iv + 1 < a.n_t; ++iv) {
        const int i = a.n_t - 1 - iv;
        float g[D];
        load_vec<D>(a.grad_h + ((int64_t)i * n_traj + idx) * D, g);
        load_vec<D>(a.h_out + ((int64_t)i * n_traj + idx) * D, y);
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] += valid ? g[d] : 0.0f;
        const int cnt = a.adj_cnt[iv];
        float s0 = gp[0];
        for (int s = 0; s + 1 < cnt; ++s) {
            const float s1 = gp[s + 1];
            fixed_adjoint_step<F, METHOD, EG>(sp, ds, s0, s1, sub_rn(s1, s0), perturb, y, lam, acc);
            s0 = s1;
        }
        gp += cnt;
    }
    if (!valid) return;
    float g0[D];
    load_vec<D>(a.grad_h + idx * D, g0);
#pragma unroll
    for (int d = 0; d < D; ++d) lam[d] += g0[d];
    if (!F::params_ok(sp)) {
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] = nanf("");
    }
    store_vec<D>(a.grad_y0 + idx * D, lam);
}

// ==============================================================================================================
// dopri5 forward (tde rk_common.py RKAdaptiveStepsizeODESolver)
//   k      storage of the 7 stage derivatives (RowsReg: registers, loops unrolled; RowsMem: shared memory, loops rolled)
//   idx    trajectory this thread integrates (clamped to a real one for padding threads)
//   valid  false for padding threads: they follow the group's control flow but contribute 0 and write nothing
//   ctrl   controller index (group or trajectory); leader writes the controller-level records
//   count  number of state elements under one controller (batch*D or D): the RMS norm is over all of them
// Comm: sum1 / sum2 (sum over the controller group, identical in every thread of the group) and any().
// ==============================================================================================================
template <int D>
HODE_HD bool any_nonfinite(const float (&v)[D]) {
    // 0 * x is NaN exactly when x is inf or NaN: one FMA chain instead of D classifications
    float z = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) z = fmaf(0.0f, v[d], z);
    return z != z;
}

// Stage inputs of one attempt: Y_i = y0 + sum_{j <= i} k_j (beta[i][j] dt), j ascending (tde: `y0 + k[..., :i+1] @ (beta_i*dt)`).
// Rows 0 .. i-1 come from storage, row i (`klast`) is still in registers from the evaluation that produced it.
template <int D, bool UNROLL, class KS>
HODE_HD void d5_stage_input(const KS& k, int i, float dtf, const float (&y0)[D], const float (&klast)[D], float (&yi)[D],
                            int base = 0) {
    float acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.0f;
    range_up<UNROLL, 6>(0, i, [&](int j) {
        float row[D];
        k.load(base + j, row);
        v_axpy<D>(acc, mul_rn(d5_beta(i, j), dtf), row, acc);
    });
    v_axpy<D>(acc, mul_rn(d5_beta(i, i), dtf), klast, acc);
    v_add<D>(yi, y0, acc);
}
// same combination with every row read from storage: y0 + sum_{j <= i} rows[base + j] (beta[i][j] dt)
template <int D, bool UNROLL, class KS>
HODE_HD void d5_stage_input_rows(const KS& k, int i, float dtf, const float (&y0)[D], float (&yi)[D], int base) {
    float acc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.0f;
    range_up<UNROLL, 6>(0, i + 1, [&](int j) {
        float row[D];
        k.load(base + j, row);
        v_axpy<D>(acc, mul_rn(d5_beta(i, j), dtf), row, acc);
    });
    v_add<D>(yi, y0, acc);
}
HODE_HD float d5_stage_time(int i, float t0f, float dtf, float t1f) {
    const float al = d5_alpha(i);
    return (al == 1.0f) ? t_prev(t1f) : add_rn(t0f, mul_rn(al, dtf));
}

// ROLLED: stage loops rolled (needs run-time row indices: RowsMem) instead of unrolled
template <class F, bool ROLLED, class PS, class Dose, class Comm, class KS>
HODE_HD void dopri5_fwd_traj(const SolveArgs& a, Comm& cm, PS sp, const Dose& ds, KS& k, int64_t idx,
                             bool valid, int64_t ctrl, bool leader, float count) {
    constexpr int D = F::D;
    constexpr bool UNROLL = !ROLLED;
    static_assert(UNROLL || KS::kDynamic, "rolled stage loops need rows that can be indexed at run time");
    static_assert(D % 2 == 0, "packed stage combination assumes even D");
    const int64_t n_traj = a.n_groups * a.batch;
    const float rtol = a.rtol_f, atol = a.atol_f;
    const float inv_count = 1.0f / count;
    const float safety = (float)a.safety, ifactor = (float)a.ifactor, dfactor = (float)a.dfactor;

    float y0[D], y1[D], klast[D];
    load_vec<D>(a.y0 + idx * D, y0);
    if (valid) store_vec<D>(a.h_out + idx * D, y0);
    const bool poisoned = !F::params_ok(sp);  // kernel variant and parameters disagree: HODE_SOLVE_NONFINITE
    if (poisoned) {
#pragma unroll
        for (int d = 0; d < D; ++d) y0[d] = nanf("");
    }

    double t0 = a.t_eval_d[0];
    const float t0f_init = (float)t0;
    F::eval(sp, t0f_init, ds, y0, klast);  // f0 = func(t[0], y0)
    k.store(0, klast);

    // ---- _select_initial_step (all float32) ----------------------------------------------------------------
    double dt;
    if (a.first_step > 0.0) {
        dt = a.first_step;
    } else {
        float scale[D];
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            scale[d] = atol + fabsf(y0[d]) * rtol;
            const float q0 = y0[d] / scale[d], q1 = klast[d] / scale[d];
            s0 += q0 * q0;
            s1 += q1 * q1;
        }
        if (!valid) { s0 = 0.0f; s1 = 0.0f; }
        cm.sum2(s0, s1);
        const float d0 = sqrtf(s0 / count), d1 = sqrtf(s1 / count);
        float h0;
        if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f;
        else h0 = (0.01f * d0) / d1;
        float f1[D];
#pragma unroll
        for (int d = 0; d < D; ++d) y1[d] = y0[d] + h0 * klast[d];
        F::eval(sp, add_rn(t0f_init, h0), ds, y1, f1);
        float s2 = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float q = (f1[d] - klast[d]) / scale[d];
            s2 += q * q;
        }
        if (!valid) s2 = 0.0f;
        cm.sum1(s2);
        const float d2 = sqrtf(s2 / count) / h0;
        float h1;
        if (d1 <= 1e-15f && d2 <= 1e-15f) {
            h1 = fmaxf(1e-6f, h0 * 1e-3f);
        } else {
            const float dm = (d2 > d1) ? d2 : d1;  // python max(d1, d2)
            h1 = powf(0.01f / dm, 0.2f);
        }
        dt = (double)nan_minf(100.0f * h0, h1);  // torch.min propagates NaN
    }

    int j = 1, nacc = 0, nrej = 0, status = HODE_SOLVE_OK;
    // int32 counters (the caps are clamped: torchdiffeq's default max_num_steps is 2**31 - 1)
    const int max_steps = (int)(a.max_num_steps < 0x7fffffffLL ? a.max_num_steps : 0x7fffffffLL);
    const int attempt_cap = (int)(a.attempt_cap < 0x7fffffffLL ? a.attempt_cap : 0x7fffffffLL);
    int n_steps = 0, attempts = 0;
    bool y0_bad = cm.any(valid && any_nonfinite<D>(y0));

    // One trip = one attempt of _adaptive_step inside _advance(t[j]).  Returns false when this controller has finished
    // (all outputs emitted, or a failure status).
    // Lock-step groups (Comm::kLockstep: several controllers in one CTA meet at one barrier per attempt): a controller that has
    // finished (`idle`) or fails here still walks the stages and the reduction, with every side effect suppressed.
    auto attempt = [&](bool idle) -> bool {
        if (!idle) {
            int fail = HODE_SOLVE_OK;
            if (poisoned) fail = HODE_SOLVE_NONFINITE;
            else if (n_steps >= max_steps || attempts >= attempt_cap) fail = HODE_SOLVE_MAX_STEPS;
            else if (!(t0 + dt > t0)) fail = HODE_SOLVE_DT_UNDERFLOW;
            else if (y0_bad) fail = HODE_SOLVE_NONFINITE;
            if (fail != HODE_SOLVE_OK) {
                status = fail;
                if constexpr (!Comm::kLockstep) return false;
                idle = true;
            }
        }
        const double t1 = t0 + dt;
        const float t0f = (float)t0, dtf = (float)dt, t1f = (float)t1;
        // ---- the 6 new stages (row 0 holds f0: FSAL); the error estimate k @ (dt c_error) is accumulated on the fly,
        //      in stage order, while each k_i is still in registers --------------------------------------------------
        float err[D];
#pragma unroll
        for (int d = 0; d < D; ++d) err[d] = 0.0f;
        k.load(0, klast);
        stage_up<UNROLL, 0, 6>([&](auto il) {
            float ti;
            stage_switch<ROLLED, 6>(il, [&](auto ic) {  // ic: compile-time stage index in both variants
                const int i = ic;
                d5_stage_input<D, true>(k, i, dtf, y0, klast, y1);
                v_axpy<D>(err, mul_rn(dtf, d5_cerr(i)), klast, err);
                ti = d5_stage_time(i, t0f, dtf, t1f);
            });
            F::eval(sp, ti, ds, y1, klast);
            k.store((int)il + 1, klast);
        });
        v_axpy<D>(err, mul_rn(dtf, d5_cerr(6)), klast, err);  // y1 = last stage input, klast = k7 = f1
        // ---- error ratio: rms(err / (atol + rtol max(|y0|, |y1|))) over the controller group.  torch.max propagates
        //      NaN: a NaN in y1 poisons the tolerance, hence the ratio, hence the step is rejected like in tde and dt
        //      becomes NaN ('underflow in dt nan').  tde additionally asserts isfinite(y0) at the start of every attempt;
        //      after the first attempt that can only trigger for a state that overflowed to +-inf WITHOUT producing a NaN
        //      in the error estimate (|y| ~ 1e38), which here ends as 'underflow in dt' a few attempts later -- the
        //      group-wide test per accepted step was ~100 instructions of warp votes for that one message -------------
        float ss = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float tol = fmaf(rtol, nan_maxf(fabsf(y0[d]), fabsf(y1[d])), atol);
            const float q = div_tol(err[d], tol);
            ss = fmaf(q, q, ss);
        }
        if (!valid) ss = 0.0f;
        if constexpr (Comm::kLockstep) {
            cm.sum1_vote(ss, idle);
            if (idle) return false;
        } else {
            cm.sum1(ss);
        }
        const float ratio = fast_sqrt(ss * inv_count);
        const bool accept = ratio <= 1.0f;
        ++attempts; ++n_steps;
        if (accept) {
            if (a.tape_y != nullptr) {
                if (nacc >= a.tape_cap) { status = HODE_SOLVE_TAPE_FULL; return false; }
                if (valid) store_vec<D>(a.tape_y + ((int64_t)nacc * n_traj + idx) * D, y0);
                if (leader) {
                    a.tape_t[(ctrl * a.tape_cap + nacc) * 2] = t0;
                    a.tape_t[(ctrl * a.tape_cap + nacc) * 2 + 1] = dt;
                }
            }
            ++nacc;
            if (a.t_eval_d[j] <= t1) {
                // _interp_fit: y_mid = y0 + k @ (dt c_mid)
                float ca[D], cb[D], cc[D], cd[D], f0[D];
                {
                    float m[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) m[d] = 0.0f;
                    range_up<UNROLL, 7>(0, 7, [&](int i) {
                        float row[D];
                        k.load(i, row);
                        const float c = mul_rn(dtf, d5_cmid(i));
#pragma unroll
                        for (int d = 0; d < D; ++d) m[d] = fmaf(row[d], c, m[d]);
                    });
                    k.load(0, f0);
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float ymid = y0[d] + m[d];
                        const float f1 = klast[d];
                        ca[d] = 2.0f * dtf * (f1 - f0[d]) - 8.0f * (y1[d] + y0[d]) + 16.0f * ymid;
                        cb[d] = dtf * (5.0f * f0[d] - 3.0f * f1) + 18.0f * y0[d] + 14.0f * y1[d] - 32.0f * ymid;
                        cc[d] = dtf * (f1 - 4.0f * f0[d]) - 11.0f * y0[d] - 5.0f * y1[d] + 16.0f * ymid;
                        cd[d] = dtf * f0[d];
                    }
                }
                while (j < a.n_t && a.t_eval_d[j] <= t1) {
                    // _interp_evaluate
                    const float x = (float)((a.t_eval_d[j] - t0) / (t1 - t0));
                    const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
                    float v[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        float tot = y0[d] + x * cd[d];
                        tot = tot + x2 * cc[d];
                        tot = tot + x3 * cb[d];
                        tot = tot + x4 * ca[d];
                        v[d] = tot;
                    }
                    if (valid) store_vec<D>(a.h_out + ((int64_t)j * n_traj + idx) * D, v);
                    ++j;
                    n_steps = 0;
                }
            }
            k.store(0, klast);  // FSAL: f0 of the next step is f1 of this one
#pragma unroll
            for (int d = 0; d < D; ++d) y0[d] = y1[d];
            t0 = t1;
        } else {
            ++nrej;
        }
        dt = optimal_step(dt, ratio, safety, ifactor, dfactor);
        return j < a.n_t;
    };
    // Controllers that share a warp (lane segments) are re-converged before every attempt, so that they walk the stage loop
    // in lock-step and share its instruction issue; a finished controller idles until the last one of its warp is done.
    bool done = !(j < a.n_t);
    if constexpr (Comm::kLockstep) {
        if (!valid) done = true;  // padding threads of a packed CTA belong to no controller
        do {  // the vote inside the attempt's barrier tells whether ANY thread of the CTA still had work in it
            const bool r = attempt(done);
            if (!done) done = !r;
        } while (!cm.all_done(done));
    } else {
        while (!cm.all_done(done)) {
            if (!done) done = !attempt(false);
        }
    }
    if (leader) {
        hode_stats st;
        st.accepted = nacc; st.rejected = nrej; st.nfe = 2 + 6 * (nacc + nrej); st.status = status;
        a.stats[ctrl] = st;
    }
}

// ==============================================================================================================
// dopri5 reverse sweep over the tape (discrete adjoint of the accepted-step map with constant step sizes, through
// FSAL and the quartic dense output).  SURVEY.md Appendix D.4, reorganised so that a step needs 9 rows of D floats
// instead of 7 stage derivatives + 7 stage adjoints:
//   rows 0..5  k_1..k_6 of the step (recomputed from the tape entry); row i is OVERWRITTEN by G_i = J(Y_i)^T kbar_i once
//              stage i has been reversed -- the stage adjoints kbar_i are never stored, they are rebuilt when needed as
//              kbar_i = dt c_mid[i] YM + dt sum_{m > i} beta[m-1][i] G_m  (m descending = the order in which the literal
//              recurrence adds them), with G_6 = adjoint of y1;
//   row 6      lambda (adjoint of y_{n+1}) on entry, G_6 after the last stage has been reversed;
//   row 7      YM = adjoint of y_mid (dense outputs emitted by this step; zero otherwise);
//   row 8      phi = adjoint of the next step's k_1, which IS this step's k_7 (FSAL).
// `nloop` >= the number of accepted steps of this trajectory's controller: the loop is END-aligned (a thread with fewer steps
// idles first with zero adjoints), so that n == 0 -- the one step with an eighth VJP -- is reached by all threads of a warp
// together and every VJP call site is warp-converged (required by the cooperative accumulators).
// ==============================================================================================================
template <class F, bool EG, bool ROLLED, class PS, class Dose, class KS, class ACC>
HODE_HD void dopri5_bwd_traj(const SolveArgs& a, PS sp, const Dose& ds, KS& R, int64_t idx, int64_t ctrl, ACC acc,
                             bool valid, int nloop) {
    constexpr int D = F::D;
    constexpr bool UNROLL = !ROLLED;
    static_assert(UNROLL || KS::kDynamic, "rolled stage loops need rows that can be indexed at run time");
    const int64_t n_traj = a.n_groups * a.batch;
    const int nacc = valid ? a.stats[ctrl].accepted : 0;
    float zero[D];
#pragma unroll
    for (int d = 0; d < D; ++d) zero[d] = 0.0f;
    R.store(6, zero);
    R.store(8, zero);
    int j = a.n_t - 1;
    // The tape entry (state, t0, dt) of step n-1 is requested while step n is being reversed: with 2-3 warps per scheduler
    // nothing else hides one HBM round trip per step (ncu: stall_long_scoreboard 1.0 per issue without it).
    constexpr bool PF = TapePrefetchD5<F>::value;
    float ynext[D];
    double t0next = 0.0, dtnext = 1.0;
    auto fetch = [&](int n, float (&y)[D], double& t0v, double& dtv) {
        if (n >= 0 && n < nacc) {
            t0v = a.tape_t[(ctrl * a.tape_cap + n) * 2];
            dtv = a.tape_t[(ctrl * a.tape_cap + n) * 2 + 1];
            load_vec<D>(a.tape_y + ((int64_t)n * n_traj + idx) * D, y);
        } else {  // idle thread: any finite state will do (its adjoints are zero); 1 keeps pow / log of the Hill terms finite
            t0v = 0.0; dtv = 1.0;
#pragma unroll
            for (int d = 0; d < D; ++d) y[d] = 1.0f;
        }
    };
    if (PF) fetch(nloop - 1, ynext, t0next, dtnext);
    for (int n = nloop - 1; n >= 0; --n) {
        const bool act = n < nacc;
        double t0, dt;
        float y0[D], klast[D], Yi[D];
        if (PF) {
            t0 = t0next; dt = dtnext;
#pragma unroll
            for (int d = 0; d < D; ++d) y0[d] = ynext[d];
            fetch(n - 1, ynext, t0next, dtnext);
        } else {
            fetch(n, y0, t0, dt);
        }
        const double t1 = t0 + dt;
        const float t0f = (float)t0, dtf = (float)dt, t1f = (float)t1;
        // FSAL: k1 of step n is k7 of step n-1 = f(prev(t1_{n-1}), y1_{n-1}); step 0 uses f(t[0], y0)
        F::eval(sp, n == 0 ? t0f : t_prev(t0f), ds, y0, klast);
        R.store(0, klast);
        stage_up<UNROLL, 0, 5>([&](auto il) {  // k_2 .. k_6 (k_7 is only needed inside its own VJP, which recomputes it)
            float ti;
            stage_switch<ROLLED, 5>(il, [&](auto ic) {
                const int i = ic;
                d5_stage_input<D, true>(R, i, dtf, y0, klast, Yi);
                ti = d5_stage_time(i, t0f, dtf, t1f);
            });
            F::eval(sp, ti, ds, Yi, klast);
            R.store((int)il + 1, klast);
        });
        // ---- dense outputs emitted by this step (t0 < t_eval[j] <= t1): adjoints of y_mid, y1 and k_7 now; those of y0
        //      and k_1 are recomputed from grad_h at the end of the step (rare, and it keeps 2 D registers free) ---------
        const int j_hi = j;
        {
            float ym[D];
#pragma unroll
            for (int d = 0; d < D; ++d) ym[d] = 0.0f;
            if (act && j >= 1 && a.t_eval_d[j] > t0) {
                float e1[D], e6[D], g[D];
#pragma unroll
                for (int d = 0; d < D; ++d) { e1[d] = 0.0f; e6[d] = 0.0f; }
                while (j >= 1 && a.t_eval_d[j] > t0) {
                    const float x = (float)((a.t_eval_d[j] - t0) / (t1 - t0));
                    const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
                    load_vec<D>(a.grad_h + ((int64_t)j * n_traj + idx) * D, g);
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float cb = x2 * g[d], bb = x3 * g[d], ab = x4 * g[d];
                        ym[d] += 16.0f * ab - 32.0f * bb + 16.0f * cb;
                        e1[d] += -8.0f * ab + 14.0f * bb - 5.0f * cb;
                        e6[d] += dtf * (2.0f * ab - 3.0f * bb + cb);
                    }
                    --j;
                }
                float r[D];
                R.load(6, r);
                v_add<D>(r, r, e1);
                R.store(6, r);
                R.load(8, r);
                v_add<D>(r, r, e6);
                R.store(8, r);
            }
            R.store(7, ym);
        }
        // ---- stages k_7 .. k_2 in reverse: i = 6 is k_7 = f(prev(t1), y1), evaluated at y1 = Y_6 ----------------------
        stage_down<UNROLL, 6, 1>([&](auto il) {
            float cot[D], row[D], g[D], ti;
            stage_switch<ROLLED, 7>(il, [&](auto ic) {
                const int i = ic;
                if (i >= 1) {
                    {
                        float acc_y[D];
#pragma unroll
                        for (int d = 0; d < D; ++d) acc_y[d] = 0.0f;
                        range_up<true, 6>(0, i, [&](int jj) {
                            R.load(jj, row);
                            v_axpy<D>(acc_y, mul_rn(d5_beta(i - 1, jj), dtf), row, acc_y);
                        });
                        v_add<D>(Yi, y0, acc_y);
                    }
                    R.load(7, row);
                    v_scale<D>(cot, mul_rn(dtf, d5_cmid(i)), row);
                    range_down<true, 7>(i + 1, 7, [&](int m) {
                        R.load(m, row);
                        v_axpy<D>(cot, mul_rn(d5_beta(m - 1, i), dtf), row, cot);
                    });
                    if (i == 6) {
                        R.load(8, row);
                        v_add<D>(cot, cot, row);
                    }
                    ti = d5_stage_time(i - 1, t0f, dtf, t1f);
                }
            });
            const int i = il;
            R.load(i, row);  // k_{i+1} for i < 6; lambda (+ dense-output terms) for i = 6
            F::template vjp<EG>(sp, ti, ds, Yi, i == 6 ? nullptr : (const float*)row, cot, g, acc);
            if (i == 6) v_add<D>(g, g, row);
            R.store(i, g);
        });
        // ---- adjoint of k_1 and of y0 ---------------------------------------------------------------------------------
        float kb0[D], yb0[D];
        {
            float row[D];
            R.load(7, row);
            v_scale<D>(kb0, mul_rn(dtf, d5_cmid(0)), row);
#pragma unroll
            for (int d = 0; d < D; ++d) yb0[d] = 0.0f;
            range_down<UNROLL, 7>(1, 7, [&](int m) {
                R.load(m, row);
                v_axpy<D>(kb0, mul_rn(d5_beta(m - 1, 0), dtf), row, kb0);
                v_add<D>(yb0, yb0, row);
            });
        }
        if (j != j_hi) {  // dense-output terms of y0 and k_1 (e = y0, d = dt f0 and their share of c, b, a)
            float g[D];
            for (int jo = j_hi; jo > j; --jo) {
                const float x = (float)((a.t_eval_d[jo] - t0) / (t1 - t0));
                const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
                load_vec<D>(a.grad_h + ((int64_t)jo * n_traj + idx) * D, g);
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float db = x * g[d], cb = x2 * g[d], bb = x3 * g[d], ab = x4 * g[d];
                    yb0[d] += g[d] - 8.0f * ab + 18.0f * bb - 11.0f * cb + (16.0f * ab - 32.0f * bb + 16.0f * cb);
                    kb0[d] += dtf * (-2.0f * ab + 5.0f * bb - 4.0f * cb + db);
                }
            }
        }
        if (n == 0) {
            float row[D], g[D];
            R.load(0, row);
            F::template vjp<EG>(sp, t0f, ds, y0, (const float*)row, kb0, g, acc);
            v_add<D>(yb0, yb0, g);
        }
        if (!act) {
#pragma unroll
            for (int d = 0; d < D; ++d) { yb0[d] = 0.0f; kb0[d] = 0.0f; }
        }
        R.store(6, yb0);  // lambda of the previous step
        R.store(8, kb0);  // phi of the previous step
    }
    if (!valid) return;
    float lam[D], g0[D];
    R.load(6, lam);
    load_vec<D>(a.grad_h + idx * D, g0);
#pragma unroll
    for (int d = 0; d < D; ++d) lam[d] += g0[d];
    if (!F::params_ok(sp)) {
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] = nanf("");
    }
    store_vec<D>(a.grad_y0 + idx * D, lam);
}


// ==============================================================================================================
// Adaptive continuous adjoint (tde adjoint.py OdeintAdjointMethod.backward with method dopri5): for every output interval
// i = n_t-1 .. 1 the augmented state (y, a, g_theta) is integrated with the dopri5 controller from t[i] back to t[i-1] in
// negated time s = -t (tde _ReverseFunc):  dy/ds = -f(-s, y),  da/ds = +J^T a,  dg/ds = +(df/dtheta)^T a;  then
// a += grad_h[i-1] and y is reset to the forward solution h[i-1].  O(1) memory: no tape.
// Error control: the 'seminorm' of tde (adjoint_options={'norm': 'seminorm'}): max(rms(err_y / tol_y), rms(err_a / tol_a)) over
// the controller group -- the parameter part g_theta is integrated (same weights, same dense output) but takes no part in
// the step-size decisions, so its stage values are only needed for ACCEPTED steps: they are formed in a second pass over the
// stored stage rows (df/dtheta^T is linear in its cotangent: stage m is called with (w_m ds) A_m, w = the 5th-order weights,
// or the dense-output weights w_m(x) for the last step of an interval, which tde interpolates back to t[i-1]).
//   rows 0..6   f(Y_m) of the attempt's stages (positive sign; dy/ds = -f is applied in the combinations)
//   rows 7..13  J(Y_m)^T A_m
// ==============================================================================================================
HODE_HD float d5_dense_weight(int m, float x) {
    // y(x) = y0 + dt sum_m w_m(x) k_m for tde's quartic (_interp_fit / _interp_evaluate); w_m(1) = c_sol[m]
    const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
    const float bm = m < 6 ? d5_beta(5, m) : 0.0f;
    float w = bm * (-5.0f * x2 + 14.0f * x3 - 8.0f * x4) + d5_cmid(m) * (16.0f * x2 - 32.0f * x3 + 16.0f * x4);
    if (m == 0) w += x - 4.0f * x2 + 5.0f * x3 - 2.0f * x4;
    if (m == 6) w += x2 - 3.0f * x3 + 2.0f * x4;
    return w;
}

template <class F, bool EG, bool ROLLED, class PS, class Dose, class Comm, class KS, class ACC>
HODE_HD void dopri5_adj_traj(const SolveArgs& a, Comm& cm, PS sp, const Dose& ds, KS& R, int64_t idx, bool valid, int64_t ctrl,
                             bool leader, float count, ACC acc) {
    constexpr int D = F::D;
    constexpr bool UNROLL = !ROLLED;
    static_assert(UNROLL || KS::kDynamic, "rolled stage loops need rows that can be indexed at run time");
    const int64_t n_traj = a.n_groups * a.batch;
    const float rtol = a.rtol_f, atol = a.atol_f;
    const float inv_count = 1.0f / count;
    const float safety = (float)a.safety, ifactor = (float)a.ifactor, dfactor = (float)a.dfactor;
    const int max_steps = (int)(a.max_num_steps < 0x7fffffffLL ? a.max_num_steps : 0x7fffffffLL);
    const int attempt_cap = (int)(a.attempt_cap < 0x7fffffffLL ? a.attempt_cap : 0x7fffffffLL);
    int nacc = 0, nrej = 0, status = HODE_SOLVE_OK, attempts = 0;
    float lam[D], y[D];
#pragma unroll
    for (int d = 0; d < D; ++d) lam[d] = 0.0f;
    const bool poisoned = !F::params_ok(sp);

    for (int iv = 0; iv + 1 < a.n_t && status == HODE_SOLVE_OK; ++iv) {
        const int i = a.n_t - 1 - iv;
        {
            float g[D];
            load_vec<D>(a.grad_h + ((int64_t)i * n_traj + idx) * D, g);
            load_vec<D>(a.h_out + ((int64_t)i * n_traj + idx) * D, y);
#pragma unroll
            for (int d = 0; d < D; ++d) lam[d] += valid ? g[d] : 0.0f;
        }
        double s0 = -a.t_eval_d[i];
        const double s_end = -a.t_eval_d[i - 1];
        float flast[D], glast[D];
        // f0 of the interval's solve: func(s0, state), no perturbation
        F::eval(sp, -(float)s0, ds, y, flast);
        vjp_state_only<F>(sp, -(float)s0, ds, y, (const float*)flast, lam, glast, acc);
        R.store(0, flast);
        R.store(7, glast);
        // ---- _select_initial_step on the augmented state with the seminorm (dy/ds = -f: signs cancel inside the norms) ----
        double dt;
        if (a.first_step > 0.0) {
            dt = a.first_step;
        } else {
            float sy[D], sa[D], q0 = 0.0f, q1 = 0.0f, r0 = 0.0f, r1 = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                sy[d] = atol + fabsf(y[d]) * rtol;
                sa[d] = atol + fabsf(lam[d]) * rtol;
                const float u0 = y[d] / sy[d], u1 = flast[d] / sy[d], v0 = lam[d] / sa[d], v1 = glast[d] / sa[d];
                q0 += u0 * u0; q1 += u1 * u1; r0 += v0 * v0; r1 += v1 * v1;
            }
            if (!valid) { q0 = q1 = r0 = r1 = 0.0f; }
            cm.sum2(q0, q1);
            cm.sum2(r0, r1);
            const float d0 = nan_maxf(sqrtf(q0 / count), sqrtf(r0 / count)), d1 = nan_maxf(sqrtf(q1 / count), sqrtf(r1 / count));
            float h0;
            if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f;
            else h0 = (0.01f * d0) / d1;
            float y1[D], a1[D], f1[D], g1[D];
#pragma unroll
            for (int d = 0; d < D; ++d) { y1[d] = y[d] - h0 * flast[d]; a1[d] = lam[d] + h0 * glast[d]; }
            const float th = -add_rn((float)s0, h0);
            F::eval(sp, th, ds, y1, f1);
            vjp_state_only<F>(sp, th, ds, y1, (const float*)f1, a1, g1, acc);
            float q2 = 0.0f, r2 = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float u = (f1[d] - flast[d]) / sy[d], v = (g1[d] - glast[d]) / sa[d];
                q2 += u * u; r2 += v * v;
            }
            if (!valid) { q2 = r2 = 0.0f; }
            cm.sum2(q2, r2);
            const float d2 = nan_maxf(sqrtf(q2 / count), sqrtf(r2 / count)) / h0;
            float h1;
            if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
            else h1 = powf(0.01f / ((d2 > d1) ? d2 : d1), 0.2f);
            dt = (double)nan_minf(100.0f * h0, h1);
        }
        int n_steps = 0;
        bool first = true;  // first step of this interval: k_1 was evaluated at s0 itself, later ones at prev(s1) (FSAL)
        bool done = false;
        while (!done) {
            if (poisoned) { status = HODE_SOLVE_NONFINITE; break; }
            if (n_steps >= max_steps || attempts >= attempt_cap) { status = HODE_SOLVE_MAX_STEPS; break; }
            if (!(s0 + dt > s0)) { status = HODE_SOLVE_DT_UNDERFLOW; break; }
            const double s1 = s0 + dt;
            const float s0f = (float)s0, dsf = (float)dt, s1f = (float)s1;
            float y1[D], a1[D], ey[D], ea[D];
#pragma unroll
            for (int d = 0; d < D; ++d) { ey[d] = 0.0f; ea[d] = 0.0f; }
            R.load(0, flast);
            R.load(7, glast);
            stage_up<UNROLL, 0, 6>([&](auto il) {
                float ts;
                stage_switch<ROLLED, 6>(il, [&](auto ic) {
                    const int m = ic;
                    d5_stage_input<D, true>(R, m, -dsf, y, flast, y1, 0);
                    d5_stage_input<D, true>(R, m, dsf, lam, glast, a1, 7);
                    const float ce = mul_rn(dsf, d5_cerr(m));
                    v_axpy<D>(ey, ce, flast, ey);
                    v_axpy<D>(ea, ce, glast, ea);
                    ts = d5_stage_time(m, s0f, dsf, s1f);
                });
                F::eval(sp, -ts, ds, y1, flast);
                vjp_state_only<F>(sp, -ts, ds, y1, (const float*)flast, a1, glast, acc);
                R.store((int)il + 1, flast);
                R.store((int)il + 8, glast);
            });
            {
                const float ce = mul_rn(dsf, d5_cerr(6));
                v_axpy<D>(ey, ce, flast, ey);
                v_axpy<D>(ea, ce, glast, ea);
            }
            float ssy = 0.0f, ssa = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float qy = div_tol(ey[d], fmaf(rtol, nan_maxf(fabsf(y[d]), fabsf(y1[d])), atol));
                const float qa = div_tol(ea[d], fmaf(rtol, nan_maxf(fabsf(lam[d]), fabsf(a1[d])), atol));
                ssy = fmaf(qy, qy, ssy);
                ssa = fmaf(qa, qa, ssa);
            }
            if (!valid) { ssy = 0.0f; ssa = 0.0f; }
            cm.sum2(ssy, ssa);
            const float ratio = nan_maxf(fast_sqrt(ssy * inv_count), fast_sqrt(ssa * inv_count));
            const bool accept = ratio <= 1.0f;
            ++attempts; ++n_steps;
            if (accept) {
                ++nacc;
                const bool last = s_end <= s1;  // the interval's output time is reached: tde interpolates back to it
                const float x = last ? (float)((s_end - s0) / (s1 - s0)) : 1.0f;
                // ---- parameter part of the accepted step: g_theta += ds sum_m w_m (df/dtheta(Y_m))^T A_m -------------------
                stage_up<UNROLL, 0, 7>([&](auto il) {
                    float Ym[D], Am[D], cot[D], frow[D], gdead[D], ts, w;
                    stage_switch<ROLLED, 7>(il, [&](auto ic) {
                        const int m = ic;
                        w = last ? d5_dense_weight(m, x) : (m < 6 ? d5_beta(5, m) : 0.0f);
                        if (m == 0) {
#pragma unroll
                            for (int d = 0; d < D; ++d) { Ym[d] = y[d]; Am[d] = lam[d]; }
                            ts = first ? s0f : t_prev(s0f);
                        } else {
                            d5_stage_input_rows<D, true>(R, m - 1, -dsf, y, Ym, 0);
                            d5_stage_input_rows<D, true>(R, m - 1, dsf, lam, Am, 7);
                            ts = d5_stage_time(m - 1, s0f, dsf, s1f);
                        }
                    });
                    if (w != 0.0f) {  // group-uniform (the tableau has c_sol[1] = c_sol[6] = 0)
                        R.load((int)il, frow);
                        v_scale<D>(cot, mul_rn(w, dsf), Am);
                        F::template vjp<EG>(sp, -ts, ds, Ym, (const float*)frow, cot, gdead, acc);
                    }
                });
                if (last) {
                    // dense output of the adjoint at s_end (the state y is reset to the forward solution, g_theta was
                    // interpolated through its weights above)
                    float m_[D], g0[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) m_[d] = 0.0f;
                    range_up<UNROLL, 7>(0, 7, [&](int mm) {
                        float row[D];
                        R.load(7 + mm, row);
                        const float c = mul_rn(dsf, d5_cmid(mm));
#pragma unroll
                        for (int d = 0; d < D; ++d) m_[d] = fmaf(row[d], c, m_[d]);
                    });
                    R.load(7, g0);
                    const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float amid = lam[d] + m_[d], f0 = g0[d], f1 = glast[d];
                        const float ca = 2.0f * dsf * (f1 - f0) - 8.0f * (a1[d] + lam[d]) + 16.0f * amid;
                        const float cb = dsf * (5.0f * f0 - 3.0f * f1) + 18.0f * lam[d] + 14.0f * a1[d] - 32.0f * amid;
                        const float cc = dsf * (f1 - 4.0f * f0) - 11.0f * lam[d] - 5.0f * a1[d] + 16.0f * amid;
                        float tot = lam[d] + x * (dsf * f0);
                        tot = tot + x2 * cc;
                        tot = tot + x3 * cb;
                        tot = tot + x4 * ca;
                        lam[d] = tot;
                    }
                    done = true;
                } else {
                    R.store(0, flast);  // FSAL
                    R.store(7, glast);
#pragma unroll
                    for (int d = 0; d < D; ++d) { y[d] = y1[d]; lam[d] = a1[d]; }
                    s0 = s1;
                    first = false;
                }
            } else {
                ++nrej;
            }
            dt = optimal_step(dt, ratio, safety, ifactor, dfactor);
        }
    }
    if (leader) {
        hode_stats st;
        st.accepted = nacc; st.rejected = nrej; st.nfe = 0; st.status = status;
        a.stats[ctrl] = st;
    }
    if (!valid) return;
    float g0[D];
    load_vec<D>(a.grad_h + idx * D, g0);
#pragma unroll
    for (int d = 0; d < D; ++d) lam[d] += g0[d];
    if (poisoned) {
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] = nanf("");
    }
    store_vec<D>(a.grad_y0 + idx * D, lam);
}


// ==============================================================================================================
// Adaptive continuous adjoint with torchdiffeq's DEFAULT (mixed) adjoint norm -- batch-coupled controller, fields with
// per-thread parameter accumulators (RocheODE, D <= 8).
// The solver state of OdeintAdjointMethod.backward is the tuple (vjp_t, y, adj_y, *adj_params) and the step-size controller
// sees  max(|vjp_t| = 0, rms(y-part), rms(adj_y-part), rms(part of parameter tensor k) for every k)  of the scaled error --
// so every ATTEMPT needs, for every parameter, the group-summed error estimate  E = ds sum_m c_err[m] P_m  and the group-summed
// increment  S = ds sum_m c_sol[m] P_m  (P_m = (df/dtheta(Y_m))^T A_m summed over the group's trajectories), and the tolerance
// uses the accumulated parameter adjoint g of the GROUP (it carries over from interval to interval):
//   per stage: the full VJP into a scratch vector, folded into E and S with the stage's weights;
//   per attempt: one vector reduction over the group (Comm::sum_vec) -> thread `leader` forms the per-tensor norms;
//   accepted step inside an interval: g += S;  last step of an interval: g += ds sum_m w_m(x) P_m from a second pass with
//   the dense-output weights (exactly how tde interpolates the parameter components back to the output time);
//   _select_initial_step of every interval includes the parameter parts of d0, d1, d2 as well.
// `pc`: group-level state in memory every thread of the group can read: g [P], q [2 P] (reduction results), r [4] (scalars).
// ==============================================================================================================
struct ParamCtl {
    float* g;
    float* q;
    float* r;
};

// max over the parameter TENSORS of rms(v[tensor] ) where v[k] = num[k] / (atol + rtol * max(|ga[k]|, |gb[k]|)):
// the 13 expert scalars (and theta_1 / theta_2) are tensors of one element, ml_net.weight and ml_net.bias are the other two
template <class F, bool EG>
HODE_HD float param_tensor_norm(const float* num, const float* ga, const float* gb, float scale_num, float rtol, float atol) {
    float worst = 0.0f;
    auto term = [&](int k) {
        const float tol = fmaf(rtol, nan_maxf(fabsf(ga[k]), fabsf(gb[k])), atol);
        const float v = (num[k] * scale_num) / tol;
        return v * v;
    };
    if (EG) {
        for (int k = 0; k < R_NSCALAR; ++k) worst = nan_maxf(worst, sqrtf(term(k)));
        if (F::ABLATE) {
            worst = nan_maxf(worst, sqrtf(term(F::OFF_TH)));
            worst = nan_maxf(worst, sqrtf(term(F::OFF_TH + 1)));
        }
    }
    if (F::ML > 0) {
        float sw = 0.0f, sb = 0.0f;
        for (int k = 0; k < F::ML * F::D; ++k) sw += term(F::OFF_W + k);
        for (int k = 0; k < F::ML; ++k) sb += term(F::OFF_B + k);
        worst = nan_maxf(worst, sqrtf(sw / (float)(F::ML * F::D)));
        worst = nan_maxf(worst, sqrtf(sb / (float)F::ML));
    }
    return worst;
}

template <class F, bool EG, class PS, class Dose, class Comm, class KS>
HODE_HD void dopri5_adj_mixed_traj(const SolveArgs& a, Comm& cm, PS sp, const Dose& ds, KS& R, int64_t idx, bool valid,
                                   int64_t ctrl, bool leader, float count, ParamCtl pc, int tid, int nthreads) {
    constexpr int D = F::D, P = F::P;
    constexpr bool UNROLL = false;
    static_assert(KS::kDynamic, "rolled stage loops need rows that can be indexed at run time");
    const int64_t n_traj = a.n_groups * a.batch;
    const float rtol = a.rtol_f, atol = a.atol_f;
    const float inv_count = 1.0f / count;
    const float safety = (float)a.safety, ifactor = (float)a.ifactor, dfactor = (float)a.dfactor;
    const int max_steps = (int)(a.max_num_steps < 0x7fffffffLL ? a.max_num_steps : 0x7fffffffLL);
    const int attempt_cap = (int)(a.attempt_cap < 0x7fffffffLL ? a.attempt_cap : 0x7fffffffLL);
    int nacc = 0, nrej = 0, status = HODE_SOLVE_OK, attempts = 0;
    float lam[D], y[D];
#pragma unroll
    for (int d = 0; d < D; ++d) lam[d] = 0.0f;
    const bool poisoned = !F::params_ok(sp);
    for (int k = tid; k < P; k += nthreads) pc.g[k] = 0.0f;
    cm.sync();

    float ES[2 * P], pm[P];  // E = ES[0 .. P), S = ES[P .. 2 P)
    // up to D = 6 (P = 27) the three vectors stay in registers (fully unrolled loops); above, they are indexed at run time and
    // live in local memory
    constexpr bool kVecRegs = P <= 32;
    auto zero_pm = [&]() {
        if constexpr (kVecRegs) {
#pragma unroll
            for (int k = 0; k < P; ++k) pm[k] = 0.0f;
        } else {
#pragma unroll 1
            for (int k = 0; k < P; ++k) pm[k] = 0.0f;
        }
    };
    auto fold = [&](float we, float ws) {  // E += we pm, S += ws pm
        if constexpr (kVecRegs) {
#pragma unroll
            for (int k = 0; k < P; ++k) { ES[k] = fmaf(we, pm[k], ES[k]); ES[P + k] = fmaf(ws, pm[k], ES[P + k]); }
        } else {
#pragma unroll 1
            for (int k = 0; k < P; ++k) { ES[k] = fmaf(we, pm[k], ES[k]); ES[P + k] = fmaf(ws, pm[k], ES[P + k]); }
        }
    };
    auto zero_es = [&]() {
        if constexpr (kVecRegs) {
#pragma unroll
            for (int k = 0; k < 2 * P; ++k) ES[k] = 0.0f;
        } else {
#pragma unroll 1
            for (int k = 0; k < 2 * P; ++k) ES[k] = 0.0f;
        }
    };

    for (int iv = 0; iv + 1 < a.n_t && status == HODE_SOLVE_OK; ++iv) {
        const int i = a.n_t - 1 - iv;
        {
            float g[D];
            load_vec<D>(a.grad_h + ((int64_t)i * n_traj + idx) * D, g);
            load_vec<D>(a.h_out + ((int64_t)i * n_traj + idx) * D, y);
#pragma unroll
            for (int d = 0; d < D; ++d) lam[d] += valid ? g[d] : 0.0f;
        }
        double s0 = -a.t_eval_d[i];
        const double s_end = -a.t_eval_d[i - 1];
        float flast[D], glast[D];
        F::eval(sp, -(float)s0, ds, y, flast);
        zero_pm();
        F::template vjp<EG>(sp, -(float)s0, ds, y, (const float*)flast, lam, glast, (float*)pm);
        R.store(0, flast);
        R.store(7, glast);
        // ---- _select_initial_step on the augmented state with the mixed norm ------------------------------------------
        double dt;
        if (a.first_step > 0.0) {
            dt = a.first_step;
        } else {
            float sy[D], sa[D], q0 = 0.0f, q1 = 0.0f, r0 = 0.0f, r1 = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                sy[d] = atol + fabsf(y[d]) * rtol;
                sa[d] = atol + fabsf(lam[d]) * rtol;
                const float u0 = y[d] / sy[d], u1 = flast[d] / sy[d], v0 = lam[d] / sa[d], v1 = glast[d] / sa[d];
                q0 += u0 * u0; q1 += u1 * u1; r0 += v0 * v0; r1 += v1 * v1;
            }
            if (!valid) { q0 = q1 = r0 = r1 = 0.0f; }
            cm.sum2(q0, q1);
            cm.sum2(r0, r1);
            // parameter parts: d0 <- g / scale, d1 <- P(s0) / scale with scale = atol + rtol |g|
            zero_es();
            if (valid) fold(1.0f, 0.0f);
            cm.template sum_vec<2 * P>(ES, pc.q);  // q[0 .. P) = P(s0) summed over the group
            if (leader) {
                pc.r[0] = param_tensor_norm<F, EG>(pc.g, pc.g, pc.g, 1.0f, rtol, atol);
                pc.r[1] = param_tensor_norm<F, EG>(pc.q, pc.g, pc.g, 1.0f, rtol, atol);
            }
            cm.sync();
            const float d0 = nan_maxf(nan_maxf(sqrtf(q0 / count), sqrtf(r0 / count)), pc.r[0]);
            const float d1 = nan_maxf(nan_maxf(sqrtf(q1 / count), sqrtf(r1 / count)), pc.r[1]);
            float h0;
            if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f;
            else h0 = (0.01f * d0) / d1;
            float y1[D], a1[D], f1[D], g1[D];
#pragma unroll
            for (int d = 0; d < D; ++d) { y1[d] = y[d] - h0 * flast[d]; a1[d] = lam[d] + h0 * glast[d]; }
            const float th = -add_rn((float)s0, h0);
            F::eval(sp, th, ds, y1, f1);
            // P at the probe point minus P(s0): fold with weights (+1 probe, -1 start) into E
            zero_es();
            if (valid) fold(-1.0f, 0.0f);
            zero_pm();
            F::template vjp<EG>(sp, th, ds, y1, (const float*)f1, a1, g1, (float*)pm);
            if (valid) fold(1.0f, 0.0f);
            float q2 = 0.0f, r2 = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float u = (f1[d] - flast[d]) / sy[d], v = (g1[d] - glast[d]) / sa[d];
                q2 += u * u; r2 += v * v;
            }
            if (!valid) { q2 = r2 = 0.0f; }
            cm.sum2(q2, r2);
            cm.sync();  // r[0 .. 1] have been read by every thread
            cm.template sum_vec<2 * P>(ES, pc.q);
            if (leader) pc.r[2] = param_tensor_norm<F, EG>(pc.q, pc.g, pc.g, 1.0f, rtol, atol);
            cm.sync();
            const float d2 = nan_maxf(nan_maxf(sqrtf(q2 / count), sqrtf(r2 / count)), pc.r[2]) / h0;
            float h1;
            if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
            else h1 = powf(0.01f / ((d2 > d1) ? d2 : d1), 0.2f);
            dt = (double)nan_minf(100.0f * h0, h1);
        }
        int n_steps = 0;
        bool first = true;
        bool done = false;
        while (!done) {
            if (poisoned) { status = HODE_SOLVE_NONFINITE; break; }
            if (n_steps >= max_steps || attempts >= attempt_cap) { status = HODE_SOLVE_MAX_STEPS; break; }
            if (!(s0 + dt > s0)) { status = HODE_SOLVE_DT_UNDERFLOW; break; }
            const double s1 = s0 + dt;
            const float s0f = (float)s0, dsf = (float)dt, s1f = (float)s1;
            float y1[D], a1[D], ey[D], ea[D];
#pragma unroll
            for (int d = 0; d < D; ++d) { ey[d] = 0.0f; ea[d] = 0.0f; }
            R.load(0, flast);
            R.load(7, glast);
            // parameter part of stage 0 (k_1 of the step: the FSAL evaluation, recomputed for its parameter VJP)
            zero_es();
            {
                float gdead[D];
                zero_pm();
                F::template vjp<EG>(sp, -(first ? s0f : t_prev(s0f)), ds, y, (const float*)flast, lam, gdead, (float*)pm);
                fold(mul_rn(dsf, d5_cerr(0)), mul_rn(dsf, d5_beta(5, 0)));
            }
            stage_up<UNROLL, 0, 6>([&](auto il) {
                float ts, we, ws;
                stage_switch<true, 6>(il, [&](auto ic) {
                    const int m = ic;
                    d5_stage_input<D, true>(R, m, -dsf, y, flast, y1, 0);
                    d5_stage_input<D, true>(R, m, dsf, lam, glast, a1, 7);
                    const float ce = mul_rn(dsf, d5_cerr(m));
                    v_axpy<D>(ey, ce, flast, ey);
                    v_axpy<D>(ea, ce, glast, ea);
                    ts = d5_stage_time(m, s0f, dsf, s1f);
                    we = mul_rn(dsf, d5_cerr(m + 1));
                    ws = (m + 1 < 6) ? mul_rn(dsf, d5_beta(5, m + 1)) : 0.0f;
                });
                F::eval(sp, -ts, ds, y1, flast);
                zero_pm();
                F::template vjp<EG>(sp, -ts, ds, y1, (const float*)flast, a1, glast, (float*)pm);
                fold(we, ws);
                R.store((int)il + 1, flast);
                R.store((int)il + 8, glast);
            });
            {
                const float ce = mul_rn(dsf, d5_cerr(6));
                v_axpy<D>(ey, ce, flast, ey);
                v_axpy<D>(ea, ce, glast, ea);
            }
            float ssy = 0.0f, ssa = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float qy = div_tol(ey[d], fmaf(rtol, nan_maxf(fabsf(y[d]), fabsf(y1[d])), atol));
                const float qa = div_tol(ea[d], fmaf(rtol, nan_maxf(fabsf(lam[d]), fabsf(a1[d])), atol));
                ssy = fmaf(qy, qy, ssy);
                ssa = fmaf(qa, qa, ssa);
            }
            if (!valid) { ssy = 0.0f; ssa = 0.0f; zero_es(); }
            cm.sum2(ssy, ssa);
            cm.template sum_vec<2 * P>(ES, pc.q);  // q[0 .. P) = E, q[P .. 2 P) = S, summed over the group
            if (leader) {
                // tolerance of parameter k: atol + rtol max(|g_k|, |g_k + S_k|); r[3] doubles as scratch-free: g + S is formed on the fly
                float worst = 0.0f;
                {
                    // reuse q[P .. 2P) in place as g + S for the tolerance, restore nothing: S is re-derived as (g + S) - g below
                    for (int k = 0; k < P; ++k) pc.q[P + k] = pc.g[k] + pc.q[P + k];
                    worst = param_tensor_norm<F, EG>(pc.q, pc.g, pc.q + P, 1.0f, rtol, atol);
                }
                pc.r[3] = worst;
            }
            cm.sync();
            const float ratio = nan_maxf(nan_maxf(fast_sqrt(ssy * inv_count), fast_sqrt(ssa * inv_count)), pc.r[3]);
            const bool accept = ratio <= 1.0f;
#if !HODE_DEVICE_BUILD
            if (leader && getenv("HODE_TRACE")) printf("[mixed] s0 %.9g dt %.9g ratio_y %.6g ratio_a %.6g ratio_p %.6g acc %d\n", s0, dt, (double)sqrtf(ssy * inv_count), (double)sqrtf(ssa * inv_count), (double)pc.r[3], (int)accept);
#endif
            ++attempts; ++n_steps;
            if (accept) {
                ++nacc;
                const bool last = s_end <= s1;
                const float x = last ? (float)((s_end - s0) / (s1 - s0)) : 1.0f;
                if (!last) {
                    for (int k = tid; k < P; k += nthreads) pc.g[k] = pc.q[P + k];  // g <- g + S
                    R.store(0, flast);  // FSAL
                    R.store(7, glast);
#pragma unroll
                    for (int d = 0; d < D; ++d) { y[d] = y1[d]; lam[d] = a1[d]; }
                    s0 = s1;
                    first = false;
                } else {
                    // parameter components interpolated to s_end through the dense-output weights: second pass over the stages
                    zero_es();
                    stage_up<UNROLL, 0, 7>([&](auto il) {
                        float Ym[D], Am[D], frow[D], gdead[D], ts, w;
                        stage_switch<true, 7>(il, [&](auto ic) {
                            const int m = ic;
                            w = d5_dense_weight(m, x);
                            if (m == 0) {
#pragma unroll
                                for (int d = 0; d < D; ++d) { Ym[d] = y[d]; Am[d] = lam[d]; }
                                ts = first ? s0f : t_prev(s0f);
                            } else {
                                d5_stage_input_rows<D, true>(R, m - 1, -dsf, y, Ym, 0);
                                d5_stage_input_rows<D, true>(R, m - 1, dsf, lam, Am, 7);
                                ts = d5_stage_time(m - 1, s0f, dsf, s1f);
                            }
                        });
                        R.load((int)il, frow);
                        zero_pm();
                        F::template vjp<EG>(sp, -ts, ds, Ym, (const float*)frow, Am, gdead, (float*)pm);
                        fold(mul_rn(w, dsf), 0.0f);
                    });
                    if (!valid) zero_es();
                    cm.template sum_vec<2 * P>(ES, pc.q);
                    for (int k = tid; k < P; k += nthreads) pc.g[k] += pc.q[k];
                    // dense output of the state adjoint at s_end
                    float m_[D], g0[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) m_[d] = 0.0f;
                    range_up<UNROLL, 7>(0, 7, [&](int mm) {
                        float row[D];
                        R.load(7 + mm, row);
                        const float c = mul_rn(dsf, d5_cmid(mm));
#pragma unroll
                        for (int d = 0; d < D; ++d) m_[d] = fmaf(row[d], c, m_[d]);
                    });
                    R.load(7, g0);
                    const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const float amid = lam[d] + m_[d], f0 = g0[d], f1 = glast[d];
                        const float ca = 2.0f * dsf * (f1 - f0) - 8.0f * (a1[d] + lam[d]) + 16.0f * amid;
                        const float cb = dsf * (5.0f * f0 - 3.0f * f1) + 18.0f * lam[d] + 14.0f * a1[d] - 32.0f * amid;
                        const float cc = dsf * (f1 - 4.0f * f0) - 11.0f * lam[d] - 5.0f * a1[d] + 16.0f * amid;
                        float tot = lam[d] + x * (dsf * f0);
                        tot = tot + x2 * cc;
                        tot = tot + x3 * cb;
                        tot = tot + x4 * ca;
                        lam[d] = tot;
                    }
                    done = true;
                }
            } else {
                ++nrej;
            }
            cm.sync();  // g (and q) settled before the next attempt reads / overwrites them
            dt = optimal_step(dt, ratio, safety, ifactor, dfactor);
        }
    }
    if (leader) {
        hode_stats st;
        st.accepted = nacc; st.rejected = nrej; st.nfe = 0; st.status = status;
        a.stats[ctrl] = st;
    }
    if (!valid) return;
    float g0[D];
    load_vec<D>(a.grad_h + idx * D, g0);
#pragma unroll
    for (int d = 0; d < D; ++d) lam[d] += g0[d];
    if (poisoned) {
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] = nanf("");
    }
    store_vec<D>(a.grad_y0 + idx * D, lam);
}

}  // namespace hode
