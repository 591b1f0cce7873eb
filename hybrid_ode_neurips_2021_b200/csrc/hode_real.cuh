// hode_real.cuh -- the real-data (ICU cohort) vector fields of the reference and their hand-derived VJPs:
//   RocheODEReal      model.py:570-657   learned dx1 / dx2 nets + GRU-ODE latent block, dose = sum of ALL past doses
//   NeuralODEReal     model.py:717-769   Linear(Z+1,H) Tanh Linear(H,Z) Tanh on [y, cumsum(action)[int(t)]]
//   NeuralODEReal2nd  model.py:660-714   second-order variant: dy = [ml_net([y, dose]), y[:Z/2]]
// Written per trajectory like hode_core.cuh and plugged into the same fixed-grid step functions (fixed_step /
// fixed_step_vjp): DecoderReal only ever runs midpoint / rk4 / euler on them (experiments/real.sh:9-17).
//
// The hidden width H is a run-time value (run_real.py:48: int((obs + action + static) * 1.2) = 43 for the ICU data); the
// latent width Z is a template parameter (state lives in registers).
//
// Dose input.  The reference evaluates, per vector-field call, an O(T) sum over every hourly dose of the stay
// (model.py:653-657).  Here a per-launch table makes it O(1):
//   RocheODEReal:  Dose(t) = sum_{i: i+1 <= t} a_i exp(kel (i+1 - t)) = exp(kel (n - t)) S[n],  n = min(T, floor(t)),
//                  S[0] = 0, S[n] = a_{n-1} + exp(-kel) S[n-1];   d Dose / d kel = exp(kel (n - t)) ((n - t) S[n] + S1[n]),
//                  S1[n] = d S[n] / d kel = exp(-kel) (S1[n-1] - S[n-1]).
//   NeuralODEReal: dose(t) = cumsum(a)[int(t)] (0 when int(t) >= T).
// Tables are [.., T+1, n_traj] (trajectory-minor: one coalesced read per evaluation).
#pragma once
#include "hode_core.cuh"

namespace hode {

HODE_HD float sigmoid_f(float x) {
#if HODE_FAST_MATH && HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
    return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
#else
    return 1.0f / (1.0f + expf(-x));
#endif
}

struct DoseTab {
    const float* s;   // S or cumsum column of this trajectory (element n at s[n * stride])
    const float* s1;  // S1 column (RocheODEReal) or nullptr
    int64_t stride;
    int T;
};

// one patient's column(s) of the dose table.  kind 0 (RocheODEReal): tab[0][n][b] = S[n], tab[1][n][b] = S1[n], n = 0..T;
// kind 1 (NeuralODEReal / 2nd): tab[0][n][b] = cumsum(a)[n] for n < T, row T = 0.
HODE_HD void real_dose_table_column(int kind, const float* a, int64_t stride_t, int32_t T, int64_t n_traj, float kel,
                                    float* tab, int64_t b) {
    if (kind == 0) {
        const float decay = expf(-kel);
        float* S = tab;
        float* S1 = tab + (int64_t)(T + 1) * n_traj;
        float s = 0.0f, s1 = 0.0f;
        S[b] = 0.0f;
        S1[b] = 0.0f;
        for (int n = 1; n <= T; ++n) {
            s1 = decay * (s1 - s);
            s = fmaf(decay, s, a[(int64_t)(n - 1) * stride_t]);
            S[(int64_t)n * n_traj + b] = s;
            S1[(int64_t)n * n_traj + b] = s1;
        }
    } else {
        float c = 0.0f;
        for (int n = 0; n < T; ++n) {
            c += a[(int64_t)n * stride_t];
            tab[(int64_t)n * n_traj + b] = c;
        }
        tab[(int64_t)T * n_traj + b] = 0.0f;
    }
}

// packed parameter layouts ------------------------------------------------------------------------------------------
//   NeuralReal : W1 [H][Z+1], b1 [H], W2 [OUT][H], b2 [OUT]                      (OUT = Z or Z/2)
//   RocheReal  : k_immunity, kel, kel2, W1a [H][3], b1a [H], W2a [H], b2a, W1b [H][2], b1b [H], W2b [H], b2b,
//                Whh [M][M], Whz [M][M], Whr [M][M]                              (M = Z - 4; absent when Z == 4)
// The staged (shared-memory) copy is the packed vector itself.

template <int Z_, bool SECOND_>
struct NeuralReal {
    static constexpr int D = Z_;
    static constexpr int IN = Z_ + 1;
    static constexpr int OUT = SECOND_ ? Z_ / 2 : Z_;
    HODE_HD static constexpr int p_count(int H) { return H * IN + H + OUT * H + OUT; }
    template <class PS>
    HODE_HD static bool params_ok(PS) { return true; }

    HODE_HD static float dose(const DoseTab& ds, float t) {
        int n = (int)t;  // Python int(t): truncation toward zero
        if (n >= ds.T) return 0.0f;
        // cumsum(action)[int(t)] (model.py:753-760): a negative index counts from the END in Python, so int(t) = -1 -- the
        // first grid point of DecoderReal's default t = arange(t0 - 1, ...) when perturb is off -- reads the LAST row (the
        // total dose); below -T the reference raises IndexError: NaN here
        if (n < 0) n += ds.T;
        return n < 0 ? nanf("") : ds.s[(int64_t)n * ds.stride];
    }

    struct Params {
        const float* p;
        int H;
        HODE_HD const float* w1(int j) const { return p + j * IN; }
        HODE_HD float b1(int j) const { return p[H * IN + j]; }
        HODE_HD float w2(int k, int j) const { return p[H * IN + H + k * H + j]; }
        HODE_HD float b2(int k) const { return p[H * IN + H + OUT * H + k]; }
        HODE_HD int off_w1() const { return 0; }
        HODE_HD int off_b1() const { return H * IN; }
        HODE_HD int off_w2() const { return H * IN + H; }
        HODE_HD int off_b2() const { return H * IN + H + OUT * H; }
    };

    HODE_HD static void eval(Params sp, float t, const DoseTab& ds, const float (&y)[Z_], float (&dy)[Z_]) {
        float in[IN], out[OUT];
#pragma unroll
        for (int i = 0; i < Z_; ++i) in[i] = y[i];
        in[Z_] = dose(ds, t);
#pragma unroll
        for (int k = 0; k < OUT; ++k) out[k] = sp.b2(k);
        for (int j = 0; j < sp.H; ++j) {
            const float* w = sp.w1(j);
            float a = sp.b1(j);
#pragma unroll
            for (int i = 0; i < IN; ++i) a = fmaf(w[i], in[i], a);
            a = tanh_f(a);
#pragma unroll
            for (int k = 0; k < OUT; ++k) out[k] = fmaf(sp.w2(k, j), a, out[k]);
        }
#pragma unroll
        for (int k = 0; k < OUT; ++k) dy[k] = tanh_f(out[k]);
        if (SECOND_) {
#pragma unroll
            for (int i = 0; i < Z_ - OUT; ++i) dy[OUT + i] = y[i];
        }
    }

    template <bool EG>
    HODE_HD static void vjp(Params sp, float t, const DoseTab& ds, const float (&y)[Z_], const float* k,
                            const float (&l)[Z_], float (&gy)[Z_], float* acc) {
        float in[IN], u[OUT];
#pragma unroll
        for (int i = 0; i < Z_; ++i) in[i] = y[i];
        in[Z_] = dose(ds, t);
        if (k != nullptr) {
#pragma unroll
            for (int o = 0; o < OUT; ++o) u[o] = l[o] * (1.0f - k[o] * k[o]);
        } else {
            float f[Z_];
            eval(sp, t, ds, y, f);
#pragma unroll
            for (int o = 0; o < OUT; ++o) u[o] = l[o] * (1.0f - f[o] * f[o]);
        }
#pragma unroll
        for (int i = 0; i < Z_; ++i) gy[i] = 0.0f;
        if (SECOND_) {
#pragma unroll
            for (int i = 0; i < Z_ - OUT; ++i) gy[i] = l[OUT + i];
        }
#pragma unroll
        for (int o = 0; o < OUT; ++o) acc[sp.off_b2() + o] += u[o];
        for (int j = 0; j < sp.H; ++j) {
            const float* w = sp.w1(j);
            float a = sp.b1(j);
#pragma unroll
            for (int i = 0; i < IN; ++i) a = fmaf(w[i], in[i], a);
            a = tanh_f(a);
            float c = 0.0f;
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                c = fmaf(sp.w2(o, j), u[o], c);
                acc[sp.off_w2() + o * sp.H + j] = fmaf(u[o], a, acc[sp.off_w2() + o * sp.H + j]);
            }
            const float del = c * (1.0f - a * a);
#pragma unroll
            for (int i = 0; i < Z_; ++i) gy[i] = fmaf(w[i], del, gy[i]);
#pragma unroll
            for (int i = 0; i < IN; ++i) acc[sp.off_w1() + j * IN + i] = fmaf(del, in[i], acc[sp.off_w1() + j * IN + i]);
            acc[sp.off_b1() + j] += del;
        }
    }
};

template <int Z_>
struct RocheReal {
    static constexpr int D = Z_;
    static constexpr int M = Z_ - 4;
    HODE_HD static constexpr int p_count(int H) { return 3 + (3 * H + H + H + 1) + (2 * H + H + H + 1) + 3 * M * M; }
    template <class PS>
    HODE_HD static bool params_ok(PS) { return true; }

    struct Params {
        const float* p;
        int H;
        HODE_HD float k_imm() const { return p[0]; }
        HODE_HD float kel() const { return p[1]; }
        HODE_HD float kel2() const { return p[2]; }
        HODE_HD int off_a() const { return 3; }                 // W1a [H][3], b1a [H], W2a [H], b2a
        HODE_HD int off_b() const { return 3 + 5 * H + 1; }     // W1b [H][2], b1b [H], W2b [H], b2b
        HODE_HD int off_hh() const { return 3 + 9 * H + 2; }
        HODE_HD int off_hz() const { return off_hh() + M * M; }
        HODE_HD int off_hr() const { return off_hh() + 2 * M * M; }
    };

    // dose and its kel-derivative from the tables
    HODE_HD static void dose(const DoseTab& ds, float t, float kel, float& d, float& d_dkel) {
        int n = (int)floorf(t);
        n = n < 0 ? 0 : (n > ds.T ? ds.T : n);
        const float dn = (float)n - t;
        const float e = expf(kel * dn);
        const float S = ds.s[(int64_t)n * ds.stride];
        const float S1 = ds.s1 != nullptr ? ds.s1[(int64_t)n * ds.stride] : 0.0f;
        d = e * S;
        d_dkel = e * fmaf(dn, S, S1);
    }

    // one of the two small expert nets: s = tanh(b2 + sum_j W2_j tanh(b1_j + W1_j . x)),  x = y[0:NI]
    template <int NI>
    HODE_HD static float small_net(const float* __restrict__ q, int H, const float (&y)[Z_]) {
        float s = q[H * NI + 2 * H];
        for (int j = 0; j < H; ++j) {
            float a = q[H * NI + j];
#pragma unroll
            for (int i = 0; i < NI; ++i) a = fmaf(q[j * NI + i], y[i], a);
            s = fmaf(q[H * NI + H + j], tanh_f(a), s);
        }
        return tanh_f(s);
    }
    template <int NI>
    HODE_HD static void small_net_vjp(const float* __restrict__ q, int H, const float (&y)[Z_], float s, float lbar,
                                      float (&gy)[Z_], float* acc) {
        const float q1 = lbar * (1.0f - s * s);
        acc[H * NI + 2 * H] += q1;
        for (int j = 0; j < H; ++j) {
            float a = q[H * NI + j];
#pragma unroll
            for (int i = 0; i < NI; ++i) a = fmaf(q[j * NI + i], y[i], a);
            const float e = tanh_f(a);
            acc[H * NI + H + j] = fmaf(q1, e, acc[H * NI + H + j]);
            const float d = q[H * NI + H + j] * q1 * (1.0f - e * e);
            acc[H * NI + j] += d;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                gy[i] = fmaf(q[j * NI + i], d, gy[i]);
                acc[j * NI + i] = fmaf(d, y[i], acc[j * NI + i]);
            }
        }
    }

    HODE_HD static void eval(Params sp, float t, const DoseTab& ds, const float (&y)[Z_], float (&dy)[Z_]) {
        const float* p = sp.p;
        dy[0] = small_net<3>(p + sp.off_a(), sp.H, y);
        dy[1] = small_net<2>(p + sp.off_b(), sp.H, y);
        dy[2] = y[1] * sp.k_imm();
        float d, dd;
        dose(ds, t, sp.kel(), d, dd);
        dy[3] = sp.kel() * d - sp.kel2() * y[3];
        if (M > 0) {
            const float* Whh = p + sp.off_hh();
            const float* Whz = p + sp.off_hz();
            const float* Whr = p + sp.off_hr();
            float rh[M > 0 ? M : 1];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                float a = 0.0f;
#pragma unroll
                for (int j = 0; j < M; ++j) a = fmaf(Whr[i * M + j], y[4 + j], a);
                rh[i] = sigmoid_f(a) * y[4 + i];
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {
                float az = 0.0f, au = 0.0f;
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    az = fmaf(Whz[i * M + j], y[4 + j], az);
                    au = fmaf(Whh[i * M + j], rh[j], au);
                }
                dy[4 + i] = (1.0f - sigmoid_f(az)) * (tanh_f(au) - y[4 + i]);
            }
        }
    }

    template <bool EG>
    HODE_HD static void vjp(Params sp, float t, const DoseTab& ds, const float (&y)[Z_], const float* k,
                            const float (&l)[Z_], float (&gy)[Z_], float* acc) {
        const float* p = sp.p;
#pragma unroll
        for (int i = 0; i < Z_; ++i) gy[i] = 0.0f;
        const float s1 = k != nullptr ? k[0] : small_net<3>(p + sp.off_a(), sp.H, y);
        const float s2 = k != nullptr ? k[1] : small_net<2>(p + sp.off_b(), sp.H, y);
        small_net_vjp<3>(p + sp.off_a(), sp.H, y, s1, l[0], gy, acc + sp.off_a());
        small_net_vjp<2>(p + sp.off_b(), sp.H, y, s2, l[1], gy, acc + sp.off_b());
        gy[1] = fmaf(l[2], sp.k_imm(), gy[1]);
        acc[0] = fmaf(l[2], y[1], acc[0]);
        float d, dd;
        dose(ds, t, sp.kel(), d, dd);
        gy[3] = fmaf(-l[3], sp.kel2(), gy[3]);
        acc[1] = fmaf(l[3], fmaf(sp.kel(), dd, d), acc[1]);
        acc[2] = fmaf(-l[3], y[3], acc[2]);
        if (M > 0) {
            constexpr int MM = M > 0 ? M : 1;
            const float* Whh = p + sp.off_hh();
            const float* Whz = p + sp.off_hz();
            const float* Whr = p + sp.off_hr();
            float r[MM], rh[MM], z[MM], u[MM];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                float a = 0.0f;
#pragma unroll
                for (int j = 0; j < M; ++j) a = fmaf(Whr[i * M + j], y[4 + j], a);
                r[i] = sigmoid_f(a);
                rh[i] = r[i] * y[4 + i];
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {
                float az = 0.0f, au = 0.0f;
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    az = fmaf(Whz[i * M + j], y[4 + j], az);
                    au = fmaf(Whh[i * M + j], rh[j], au);
                }
                z[i] = sigmoid_f(az);
                u[i] = tanh_f(au);
            }
            float grh[MM];
#pragma unroll
            for (int j = 0; j < M; ++j) grh[j] = 0.0f;
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const float li = l[4 + i];
                const float gz = -li * (u[i] - y[4 + i]);
                const float gu = li * (1.0f - z[i]);
                gy[4 + i] = fmaf(-li, 1.0f - z[i], gy[4 + i]);
                const float gp = gu * (1.0f - u[i] * u[i]);
                const float gaz = gz * z[i] * (1.0f - z[i]);
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    grh[j] = fmaf(Whh[i * M + j], gp, grh[j]);
                    acc[sp.off_hh() + i * M + j] = fmaf(gp, rh[j], acc[sp.off_hh() + i * M + j]);
                    gy[4 + j] = fmaf(Whz[i * M + j], gaz, gy[4 + j]);
                    acc[sp.off_hz() + i * M + j] = fmaf(gaz, y[4 + j], acc[sp.off_hz() + i * M + j]);
                }
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                gy[4 + j] = fmaf(grh[j], r[j], gy[4 + j]);
                const float gar = grh[j] * y[4 + j] * r[j] * (1.0f - r[j]);
#pragma unroll
                for (int kk = 0; kk < M; ++kk) {
                    gy[4 + kk] = fmaf(Whr[j * M + kk], gar, gy[4 + kk]);
                    acc[sp.off_hr() + j * M + kk] = fmaf(gar, y[4 + kk], acc[sp.off_hr() + j * M + kk]);
                }
            }
        }
    }
};

}  // namespace hode
