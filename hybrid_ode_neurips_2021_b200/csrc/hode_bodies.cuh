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
    // continuous adjoint: `grid` holds the n_t - 1 reversed-time interval grids back to back (n_grid points in total),
    // adj_cnt[iv] = number of grid points of interval iv (iv = 0 is the LAST output interval)
    const int32_t* adj_cnt;
};

// Reverse sweeps request the tape entry of the NEXT step to be reversed one step ahead (D more live registers).  Only where
// registers allow: RocheODE with <= 64 packed parameters (D <= 8).  At D = 12 (117 accumulators, kernels already spill) the
// extra live state costs more than the hidden latency (measured: dopri5 reverse sweep 11.9 -> 13.4 ms at the C3 shape).
template <class F> struct TapePrefetch { static constexpr bool value = false; };
template <int D_, bool H_, bool A_> struct TapePrefetch<Roche<D_, H_, A_>> { static constexpr bool value = Roche<D_, H_, A_>::P <= 64; };

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
template <class F, int METHOD, class PS, class Dose>
HODE_HD void fixed_fwd_traj(const SolveArgs& a, PS sp, const Dose& ds, int64_t idx) {
    constexpr int D = F::D;
    const int64_t n_traj = a.n_groups * a.batch;
    float y[D], y1[D];
    load_vec<D>(a.y0 + idx * D, y);
    store_vec<D>(a.h_out + idx * D, y);  // solution[0] = y0
    if (!F::params_ok(sp)) {  // kernel variant and parameters disagree (hode_cfg.flags): fail loudly
#pragma unroll
        for (int d = 0; d < D; ++d) y[d] = nanf("");
    }
    int j = 1;
    const bool perturb = a.perturb != 0;
    // loop-carried scalars instead of per-step index arithmetic: grid time, tape cursor, next output time
    float t0 = a.grid[0];
    float* tp = a.tape_y != nullptr ? a.tape_y + idx * D : nullptr;
    const int64_t tape_stride = n_traj * D;
    float tj = (j < a.n_t) ? a.t_eval_f[j] : INFINITY;
    // one grid step ya -> yb, with the outputs it emits
    auto one_step = [&](int s, float (&ya)[D], float (&yb)[D]) {
        const float t1 = a.grid[s + 1];
        const float dt = sub_rn(t1, t0);
        if (tp != nullptr) {
            store_vec<D>(tp, ya);
            tp += tape_stride;
        }
        fixed_step<F, METHOD>(sp, ds, t0, t1, dt, perturb, ya, yb);
        if (t1 >= tj) {  // rare: an output time is reached (false for NaN times, like the reference's `while`)
            while (j < a.n_t && t1 >= a.t_eval_f[j]) {
                const float te = a.t_eval_f[j];
                float* o = a.h_out + ((int64_t)j * n_traj + idx) * D;
                if (te == t0) {
                    store_vec<D>(o, ya);
                } else if (te == t1) {
                    store_vec<D>(o, yb);
                } else {  // _linear_interp
                    const float slope = div_rn(sub_rn(te, t0), sub_rn(t1, t0));
                    float v[D];
#pragma unroll
                    for (int d = 0; d < D; ++d) v[d] = ya[d] + slope * (yb[d] - ya[d]);
                    store_vec<D>(o, v);
                }
                ++j;
            }
            tj = (j < a.n_t) ? a.t_eval_f[j] : INFINITY;
        }
        t0 = t1;
    };
    // (a two-steps-per-trip variant with y / y1 swapping roles was measured: no difference)
    for (int s = 0; s + 1 < a.n_grid; ++s) {
        one_step(s, y, y1);
#pragma unroll
        for (int d = 0; d < D; ++d) y[d] = y1[d];
    }
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
//   idx    trajectory this thread integrates (clamped to a real one for padding threads)
//   valid  false for padding threads: they follow the group's control flow but contribute 0 and write nothing
//   ctrl   controller index (group or trajectory); leader writes the controller-level records
//   count  number of state elements under one controller (batch*D or D): the RMS norm is over all of them
// ==============================================================================================================
template <class F, class PS, class Dose, class Comm>
HODE_HD void dopri5_fwd_traj(const SolveArgs& a, Comm& cm, PS sp, const Dose& ds, int64_t idx,
                             bool valid, int64_t ctrl, bool leader, float count) {
    constexpr int D = F::D;
    const int64_t n_traj = a.n_groups * a.batch;
    const Dopri5Tab T = dopri5_tab();
    const float rtol = a.rtol_f, atol = a.atol_f;

    float y0[D], y1[D];
    StageRegs<D> ks;
    float (&k)[7][D] = ks.v;
    load_vec<D>(a.y0 + idx * D, y0);
    if (valid) store_vec<D>(a.h_out + idx * D, y0);
    const bool poisoned = !F::params_ok(sp);  // kernel variant and parameters disagree: HODE_SOLVE_NONFINITE
    if (poisoned) {
#pragma unroll
        for (int d = 0; d < D; ++d) y0[d] = nanf("");
    }

    double t0 = a.t_eval_d[0];
    const float t0f_init = (float)t0;
    F::eval(sp, t0f_init, ds, y0, k[0]);  // f0 = func(t[0], y0)

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
            const float q0 = y0[d] / scale[d], q1 = k[0][d] / scale[d];
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
        for (int d = 0; d < D; ++d) y1[d] = y0[d] + h0 * k[0][d];
        F::eval(sp, add_rn(t0f_init, h0), ds, y1, f1);
        float s2 = 0.0f, dummy = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float q = (f1[d] - k[0][d]) / scale[d];
            s2 += q * q;
        }
        if (!valid) s2 = 0.0f;
        cm.sum2(s2, dummy);
        const float d2 = sqrtf(s2 / count) / h0;
        float h1;
        if (d1 <= 1e-15f && d2 <= 1e-15f) {
            h1 = fmaxf(1e-6f, h0 * 1e-3f);
        } else {
            const float dm = (d2 > d1) ? d2 : d1;  // python max(d1, d2)
            h1 = powf(0.01f / dm, 0.2f);
        }
        const float hh = 100.0f * h0;
        // torch.min propagates NaN
        dt = (double)((hh != hh || h1 != h1) ? (hh + h1) : (hh < h1 ? hh : h1));
    }

    int j = 1, nacc = 0, nrej = 0, status = HODE_SOLVE_OK;
    int64_t n_steps = 0, attempts = 0;
    float bad0 = 0.0f, dummy0 = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) bad0 += isfinite(y0[d]) ? 0.0f : 1.0f;
    if (!valid) bad0 = 0.0f;
    cm.sum2(bad0, dummy0);
    bool y0_bad = bad0 > 0.0f;

    while (j < a.n_t) {
        // _advance(t[j]): while next_t > rk_state.t1 -> _adaptive_step
        if (poisoned) { status = HODE_SOLVE_NONFINITE; break; }
        if (n_steps >= a.max_num_steps || attempts >= a.attempt_cap) { status = HODE_SOLVE_MAX_STEPS; break; }
        if (!(t0 + dt > t0)) { status = HODE_SOLVE_DT_UNDERFLOW; break; }
        if (y0_bad) { status = HODE_SOLVE_NONFINITE; break; }
        const double t1 = t0 + dt;
        const float t0f = (float)t0, dtf = (float)dt, t1f = (float)t1;
        dopri5_stages<F>(sp, ds, T, t0f, dtf, t1f, y0, ks, y1);
        // error estimate and ratio
        float ss = 0.0f, bad = 0.0f;
        float err[D];
#pragma unroll
        for (int d = 0; d < D; d += 2) {
            float e0 = 0.0f, e1 = 0.0f;
#pragma unroll
            for (int i = 0; i < 7; ++i) fma2s(mul_rn(dtf, T.c_err[i]), k[i][d], k[i][d + 1], e0, e1, e0, e1);
            err[d] = e0; err[d + 1] = e1;
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float e = err[d];
            const float tol = atol + rtol * fmaxf(fabsf(y0[d]), fabsf(y1[d]));
            const float q = e / tol;
            ss += q * q;
            bad += isfinite(y1[d]) ? 0.0f : 1.0f;
        }
        // fmaxf drops NaN; torch.max propagates it.  A NaN in y1 must poison the ratio like it does in the reference.
        if (bad > 0.0f) ss = nanf("");
        if (!valid) { ss = 0.0f; bad = 0.0f; }
        cm.sum2(ss, bad);
        const float ratio = fabsf(sqrtf(ss / count));
        const bool accept = ratio <= 1.0f;
        ++attempts; ++n_steps;
        if (accept) {
            if (a.tape_y != nullptr) {
                if (nacc >= a.tape_cap) { status = HODE_SOLVE_TAPE_FULL; break; }
                if (valid) store_vec<D>(a.tape_y + ((int64_t)nacc * n_traj + idx) * D, y0);
                if (leader) {
                    a.tape_t[(ctrl * a.tape_cap + nacc) * 2] = t0;
                    a.tape_t[(ctrl * a.tape_cap + nacc) * 2 + 1] = dt;
                }
            }
            ++nacc;
            if (a.t_eval_d[j] <= t1) {
                // _interp_fit
                float ca[D], cb[D], cc[D], cd[D];
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    float m = 0.0f;
#pragma unroll
                    for (int i = 0; i < 7; ++i) m = fmaf(k[i][d], mul_rn(dtf, T.c_mid[i]), m);
                    const float ymid = y0[d] + m;
                    const float f0 = k[0][d], f1 = k[6][d];
                    ca[d] = 2.0f * dtf * (f1 - f0) - 8.0f * (y1[d] + y0[d]) + 16.0f * ymid;
                    cb[d] = dtf * (5.0f * f0 - 3.0f * f1) + 18.0f * y0[d] + 14.0f * y1[d] - 32.0f * ymid;
                    cc[d] = dtf * (f1 - 4.0f * f0) - 11.0f * y0[d] - 5.0f * y1[d] + 16.0f * ymid;
                    cd[d] = dtf * f0;
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
#pragma unroll
            for (int d = 0; d < D; ++d) { y0[d] = y1[d]; k[0][d] = k[6][d]; }
            t0 = t1;
            y0_bad = bad > 0.0f;
        } else {
            ++nrej;
        }
        dt = optimal_step(dt, ratio, a.safety, a.ifactor, a.dfactor);
    }
    if (leader) {
        hode_stats st;
        st.accepted = nacc; st.rejected = nrej; st.nfe = 2 + 6 * (nacc + nrej); st.status = status;
        a.stats[ctrl] = st;
    }
}

// ==============================================================================================================
// dopri5 reverse sweep over the tape (discrete adjoint of the accepted-step map with constant step sizes, through
// FSAL and the quartic dense output).  SURVEY.md Appendix D.4.
// ==============================================================================================================
template <class F, bool EG, class PS, class Dose, class KS, class ACC>
HODE_HD void dopri5_bwd_traj(const SolveArgs& a, PS sp, const Dose& ds, int64_t idx, int64_t ctrl, ACC acc, KS& k,
                             KS& kb, bool valid = true) {
    constexpr int D = F::D;
    const int64_t n_traj = a.n_groups * a.batch;
    const Dopri5Tab T = dopri5_tab();
    const int nacc = a.stats[ctrl].accepted;
    float lam[D], phi[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { lam[d] = 0.0f; phi[d] = 0.0f; }
    int j = a.n_t - 1;
    // tape entry (state, t0, dt) of step n-1 is requested while step n is being reversed (see fixed_bwd_traj)
    constexpr bool PF = TapePrefetch<F>::value;
    float ynext[D];
    double t0next = 0.0, dtnext = 0.0;
    if (PF && nacc > 0) {
        load_vec<D>(a.tape_y + ((int64_t)(nacc - 1) * n_traj + idx) * D, ynext);
        t0next = a.tape_t[(ctrl * a.tape_cap + nacc - 1) * 2];
        dtnext = a.tape_t[(ctrl * a.tape_cap + nacc - 1) * 2 + 1];
    }
    for (int n = nacc - 1; n >= 0; --n) {
        const double t0 = PF ? t0next : a.tape_t[(ctrl * a.tape_cap + n) * 2];
        const double dt = PF ? dtnext : a.tape_t[(ctrl * a.tape_cap + n) * 2 + 1];
        const double t1 = t0 + dt;
        const float t0f = (float)t0, dtf = (float)dt, t1f = (float)t1;
        float y0[D], y1[D], yb0[D], yb1[D], g[D], kr[D], lr[D];
        if (PF) {
#pragma unroll
            for (int d = 0; d < D; ++d) y0[d] = ynext[d];
            if (n > 0) {
                load_vec<D>(a.tape_y + ((int64_t)(n - 1) * n_traj + idx) * D, ynext);
                t0next = a.tape_t[(ctrl * a.tape_cap + n - 1) * 2];
                dtnext = a.tape_t[(ctrl * a.tape_cap + n - 1) * 2 + 1];
            }
        } else {
            load_vec<D>(a.tape_y + ((int64_t)n * n_traj + idx) * D, y0);
        }
        // FSAL: k1 of step n is k7 of step n-1 = f(prev(t1_{n-1}), y1_{n-1}); step 0 uses f(t[0], y0)
        F::eval(sp, n == 0 ? t0f : t_prev(t0f), ds, y0, kr);
        stage_set_row<D>(k, 0, kr);
        dopri5_stages<F>(sp, ds, T, t0f, dtf, t1f, y0, k, y1);
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int d = 0; d < D; ++d) kb.set(i, d, 0.0f);
#pragma unroll
        for (int d = 0; d < D; ++d) { yb1[d] = lam[d]; yb0[d] = 0.0f; kb.set(6, d, phi[d]); }
        // dense outputs emitted by this step: t0 < t_eval[j] <= t1
        while (j >= 1 && a.t_eval_d[j] > t0) {
            const float x = (float)((a.t_eval_d[j] - t0) / (t1 - t0));
            const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
            load_vec<D>(a.grad_h + ((int64_t)j * n_traj + idx) * D, g);
            if (!valid) {  // padding lane of a warp-cooperative accumulator: follows the control flow, contributes 0
#pragma unroll
                for (int d = 0; d < D; ++d) g[d] = 0.0f;
            }
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float eb = g[d], db = x * g[d], cb = x2 * g[d], bb = x3 * g[d], ab = x4 * g[d];
                const float ymb = 16.0f * ab - 32.0f * bb + 16.0f * cb;
                yb0[d] += eb - 8.0f * ab + 18.0f * bb - 11.0f * cb + ymb;
                yb1[d] += -8.0f * ab + 14.0f * bb - 5.0f * cb;
                kb.set(0, d, kb.get(0, d) + dtf * (-2.0f * ab + 5.0f * bb - 4.0f * cb + db));
                kb.set(6, d, kb.get(6, d) + dtf * (2.0f * ab - 3.0f * bb + cb));
#pragma unroll
                for (int i = 0; i < 7; ++i) kb.set(i, d, fmaf(mul_rn(dtf, T.c_mid[i]), ymb, kb.get(i, d)));
            }
            --j;
        }
        // k7 = f(prev(t1), y1)
        stage_row<D>(k, 6, kr);
        stage_row<D>(kb, 6, lr);
        F::template vjp<EG>(sp, t_prev(t1f), ds, y1, kr, lr, g, acc);
#pragma unroll
        for (int d = 0; d < D; ++d) yb1[d] += g[d];
        // y1 = y0 + sum_{j<6} k_j * (beta[5][j]*dt)
#pragma unroll
        for (int d = 0; d < D; d += 2) {
            add2(yb0[d], yb0[d + 1], yb1[d], yb1[d + 1], yb0[d], yb0[d + 1]);
#pragma unroll
            for (int jj = 0; jj < 6; ++jj) {
                float o0, o1;
                fma2s(mul_rn(T.beta[5][jj], dtf), yb1[d], yb1[d + 1], kb.get(jj, d), kb.get(jj, d + 1), o0, o1);
                kb.set(jj, d, o0); kb.set(jj, d + 1, o1);
            }
        }
        // stages k6 .. k2  (k[i] = f(t_i, Y_i), Y_i = y0 + sum_{j<i} k_j * (beta[i-1][j]*dt))
#pragma unroll
        for (int i = 5; i >= 1; --i) {
            float ti;
            if (T.alpha[i - 1] == 1.0f) ti = t_prev(t1f);
            else ti = add_rn(t0f, mul_rn(T.alpha[i - 1], dtf));
            float Yi[D];
#pragma unroll
            for (int d = 0; d < D; d += 2) {
                float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
                for (int jj = 0; jj < i; ++jj) fma2s(mul_rn(T.beta[i - 1][jj], dtf), k.get(jj, d), k.get(jj, d + 1), s0, s1, s0, s1);
                add2(y0[d], y0[d + 1], s0, s1, Yi[d], Yi[d + 1]);
            }
            stage_row<D>(k, i, kr);
            stage_row<D>(kb, i, lr);
            F::template vjp<EG>(sp, ti, ds, Yi, kr, lr, g, acc);
#pragma unroll
            for (int d = 0; d < D; d += 2) {
                add2(yb0[d], yb0[d + 1], g[d], g[d + 1], yb0[d], yb0[d + 1]);
#pragma unroll
                for (int jj = 0; jj < i; ++jj) {
                    float o0, o1;
                    fma2s(mul_rn(T.beta[i - 1][jj], dtf), g[d], g[d + 1], kb.get(jj, d), kb.get(jj, d + 1), o0, o1);
                    kb.set(jj, d, o0); kb.set(jj, d + 1, o1);
                }
            }
        }
        if (n == 0) {
            stage_row<D>(k, 0, kr);
            stage_row<D>(kb, 0, lr);
            F::template vjp<EG>(sp, t0f, ds, y0, kr, lr, g, acc);
#pragma unroll
            for (int d = 0; d < D; ++d) yb0[d] += g[d];
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d) phi[d] = kb.get(0, d);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) lam[d] = yb0[d];
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

// register-resident stage storage (the default)
template <class F, bool EG, class PS, class Dose, class ACC>
HODE_HD void dopri5_bwd_traj(const SolveArgs& a, PS sp, const Dose& ds, int64_t idx, int64_t ctrl, ACC acc,
                             bool valid = true) {
    StageRegs<F::D> k, kb;
    dopri5_bwd_traj<F, EG>(a, sp, ds, idx, ctrl, acc, k, kb, valid);
}

}  // namespace hode
