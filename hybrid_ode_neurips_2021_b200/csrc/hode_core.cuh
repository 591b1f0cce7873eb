// hode_core.cuh -- per-trajectory arithmetic of the hybrid-ODE hot path: vector fields (+ hand-derived VJPs),
// fixed-grid step functions, the dopri5 attempt, dense output, and their reverse sweeps.
//
// Everything here is written per trajectory ("one thread = one patient, state in registers") and is shared by the
// sm_100a kernels in hode_launch.cuh (instantiated in inst_*.cu).  It also compiles as plain C++ (HODE_HOSTSIM) for tests/hostsim, a TEST-ONLY
// thread-emulation used to debug the derivations in a container that has no GPU; the product never loads that.
//
// Reference semantics followed (file:line into the reference; "tde" = torchdiffeq 0.2.2, restated in oracle/odeint.py):
//   RocheODE.forward / dose_at_time     model.py:509-555
//   NeuralODE.forward / dose_at_time    model.py:1015-1026
//   tde fixed_grid.py step functions, solvers.py FixedGridODESolver.integrate, rk_common.py _runge_kutta_step,
//   misc.py _select_initial_step/_compute_error_ratio/_optimal_step_size, interp.py _interp_fit/_interp_evaluate
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(HODE_HOSTSIM)
#define HODE_HD __host__ __device__ __forceinline__
#define HODE_D __device__ __forceinline__
#define HODE_DEVICE_BUILD 1
#else
#define HODE_HD inline
#define HODE_D inline
#define HODE_DEVICE_BUILD 0
#endif

#ifndef HODE_FAST_MATH
#define HODE_FAST_MATH 1
#endif

namespace hode {

// ------------------------------------------------------------------------------------------------------------
// exactly-rounded fp32 primitives for TIME arithmetic.  PyTorch evaluates `t0 + alpha*dt` as two separately rounded
// float32 ops; a contracted FMA would move stage times by an ulp and flip `t >= dose_time` (model.py:512).
// ------------------------------------------------------------------------------------------------------------
#if HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
HODE_D float mul_rn(float a, float b) { return __fmul_rn(a, b); }
HODE_D float add_rn(float a, float b) { return __fadd_rn(a, b); }
HODE_D float sub_rn(float a, float b) { return __fsub_rn(a, b); }
HODE_D float div_rn(float a, float b) { return __fdiv_rn(a, b); }
#else
// host build is compiled with -ffp-contract=off
HODE_HD float mul_rn(float a, float b) { volatile float r = a * b; return r; }
HODE_HD float add_rn(float a, float b) { volatile float r = a + b; return r; }
HODE_HD float sub_rn(float a, float b) { volatile float r = a - b; return r; }
HODE_HD float div_rn(float a, float b) { volatile float r = a / b; return r; }
#endif

// tde _nextafter(t, t+1) / _nextafter(t, t-1)
#if HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
// libm's nextafterf is ~25 instructions with three branches, paid twice per dopri5 attempt (the two stages at t0 + dt) and
// once per perturbed fixed-grid step (ncu: 8 % of the dopri5 forward kernel's instructions at D = 6).  Same function on the
// bit pattern: toward a larger value the magnitude grows for t > 0 and shrinks for t < 0; t + 1 == t (|t| >= 2^24, inf)
// returns t like nextafter(x, x); NaN propagates through t + 1.
HODE_D float t_next(float t) {
    const float y = t + 1.0f;
    if (!(y > t)) return y;
    const int b = __float_as_int(t);
    return __int_as_float(t > 0.0f ? b + 1 : (t < 0.0f ? b - 1 : 0x00000001));
}
HODE_D float t_prev(float t) {
    const float y = t - 1.0f;
    if (!(y < t)) return y;
    const int b = __float_as_int(t);
    return __int_as_float(t > 0.0f ? b - 1 : (t < 0.0f ? b + 1 : (int)0x80000001u));
}
#else
HODE_HD float t_next(float t) { return nextafterf(t, t + 1.0f); }
HODE_HD float t_prev(float t) { return nextafterf(t, t - 1.0f); }
#endif

// ---- transcendental primitives ----------------------------------------------------------------------------------
// HODE_FAST_MATH=1 (device default): MUFU-based forms with ~1e-7 absolute error (DESIGN.md "math"): the kernels are
// issue-bound, and libm's tanhf/expf/division cost 4-5x the instructions.  HODE_FAST_MATH=0 keeps the libm calls.
#if HODE_FAST_MATH && HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
HODE_D float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
HODE_D float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// tanh(x) = 1 - 2/(2^(2x log2 e) + 1): 2 MUFU + 3 FMA-pipe ops; saturates correctly at +-inf, NaN propagates
HODE_D float tanh_f(float x) { return fmaf(-2.0f, rcp_approx(ex2_approx(x * 2.8853900817779268f) + 1.0f), 1.0f); }
// exp(kel * d) with kl2 = kel * log2(e) pre-multiplied
HODE_D float exp_scaled(float kl2, float /*kel*/, float d) { return ex2_approx(kl2 * d); }
HODE_D float rcp_f(float x) { return rcp_approx(x); }
// tanh of a pre-scaled argument: xs = x * 2 log2(e).  The RocheODE kernels fold the factor into the staged ml_net
// weights, which removes one multiply per hidden unit per evaluation.
#define HODE_FOLD_TANH 1
HODE_D float tanh_pre(float xs) { return fmaf(-2.0f, rcp_approx(ex2_approx(xs) + 1.0f), 1.0f); }
#else
#define HODE_FOLD_TANH 0
HODE_HD float tanh_pre(float x) { return tanhf(x); }
HODE_HD float tanh_f(float x) { return tanhf(x); }
HODE_HD float exp_scaled(float /*kl2*/, float kel, float d) { return expf(mul_rn(kel, d)); }
HODE_HD float rcp_f(float x) { return 1.0f / x; }
#endif

#if HODE_FOLD_TANH
static constexpr float kTanhPre = 2.8853900817779268f, kTanhPreInv = 0.34657359027997264f;
#else
static constexpr float kTanhPre = 1.0f, kTanhPreInv = 1.0f;
#endif

// ------------------------------------------------------------------------------------------------------------
// Packed FP32 pairs.  sm_100a has FFMA2 / FADD2 / FMUL2 (fma.rn.f32x2 ...): two IEEE operations per lane in ONE issue slot,
// with a scalar-broadcast operand form (R.F32) and a uniform-register pair form (UR.F32x2) for constant-bank weights.
// The solver kernels are issue-bound, so the per-dimension loops are written on pairs
// (v[2i], v[2i+1]).  Every packed operation is the same correctly rounded fma / add / mul per element as the scalar code
// it replaces (the host build below IS that scalar code): results do not change, only the number of issue slots.
// ------------------------------------------------------------------------------------------------------------
#ifndef HODE_PACKED
#define HODE_PACKED 1
#endif
#if HODE_PACKED && HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
// o = s * b + c   (s broadcast)
HODE_D void fma2s(float s, float b0, float b1, float c0, float c1, float& o0, float& o1) {
    const float2 r = __ffma2_rn(make_float2(s, s), make_float2(b0, b1), make_float2(c0, c1));
    o0 = r.x; o1 = r.y;
}
HODE_D void fma2(float a0, float a1, float b0, float b1, float c0, float c1, float& o0, float& o1) {
    const float2 r = __ffma2_rn(make_float2(a0, a1), make_float2(b0, b1), make_float2(c0, c1));
    o0 = r.x; o1 = r.y;
}
HODE_D void add2(float a0, float a1, float b0, float b1, float& o0, float& o1) {
    const float2 r = __fadd2_rn(make_float2(a0, a1), make_float2(b0, b1));
    o0 = r.x; o1 = r.y;
}
HODE_D void mul2(float a0, float a1, float b0, float b1, float& o0, float& o1) {
    const float2 r = __fmul2_rn(make_float2(a0, a1), make_float2(b0, b1));
    o0 = r.x; o1 = r.y;
}
#else
HODE_HD void fma2s(float s, float b0, float b1, float c0, float c1, float& o0, float& o1) { o0 = fmaf(s, b0, c0); o1 = fmaf(s, b1, c1); }
HODE_HD void fma2(float a0, float a1, float b0, float b1, float c0, float c1, float& o0, float& o1) { o0 = fmaf(a0, b0, c0); o1 = fmaf(a1, b1, c1); }
HODE_HD void add2(float a0, float a1, float b0, float b1, float& o0, float& o1) { o0 = add_rn(a0, b0); o1 = add_rn(a1, b1); }
HODE_HD void mul2(float a0, float a1, float b0, float b1, float& o0, float& o1) { o0 = mul_rn(a0, b0); o1 = mul_rn(a1, b1); }
#endif

// element-wise vector helpers over the D state dimensions (D is even for every compiled field)
template <int D>  // o = s * a + c
HODE_HD void v_axpy(float (&o)[D], float s, const float (&a)[D], const float (&c)[D]) {
#pragma unroll
    for (int i = 0; i + 1 < D; i += 2) fma2s(s, a[i], a[i + 1], c[i], c[i + 1], o[i], o[i + 1]);
    if (D & 1) o[D - 1] = fmaf(s, a[D - 1], c[D - 1]);
}
template <int D>  // o = a + b
HODE_HD void v_add(float (&o)[D], const float (&a)[D], const float (&b)[D]) {
#pragma unroll
    for (int i = 0; i + 1 < D; i += 2) add2(a[i], a[i + 1], b[i], b[i + 1], o[i], o[i + 1]);
    if (D & 1) o[D - 1] = add_rn(a[D - 1], b[D - 1]);
}
template <int D>  // o = a - b
HODE_HD void v_sub(float (&o)[D], const float (&a)[D], const float (&b)[D]) {
#pragma unroll
    for (int i = 0; i + 1 < D; i += 2) add2(a[i], a[i + 1], -b[i], -b[i + 1], o[i], o[i + 1]);
    if (D & 1) o[D - 1] = sub_rn(a[D - 1], b[D - 1]);
}
template <int D>  // o = s * a
HODE_HD void v_scale(float (&o)[D], float s, const float (&a)[D]) {
#pragma unroll
    for (int i = 0; i + 1 < D; i += 2) mul2(s, s, a[i], a[i + 1], o[i], o[i + 1]);
    if (D & 1) o[D - 1] = mul_rn(s, a[D - 1]);
}

// two tanh_pre at once: the add and the final fma are packed
#if HODE_FAST_MATH && HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
HODE_D void tanh_pre2(float x0, float x1, float& o0, float& o1) {
    float q0, q1;
    add2(ex2_approx(x0), ex2_approx(x1), 1.0f, 1.0f, q0, q1);
    fma2s(-2.0f, rcp_approx(q0), rcp_approx(q1), 1.0f, 1.0f, o0, o1);
}
#else
HODE_HD void tanh_pre2(float x0, float x1, float& o0, float& o1) { o0 = tanh_pre(x0); o1 = tanh_pre(x1); }
#endif

// x ** p with a float32 tensor exponent (model.py:529, 537-538).  p == 2 (RochConfig default) is a multiply.
HODE_HD float pow_hill(float x, float p) { return (p == 2.0f) ? x * x : powf(x, p); }
// d/dx x**p = p * x**(p-1)   (autograd pow_backward_self; zero where p == 0)
HODE_HD float dpow_hill(float x, float p) {
    if (p == 2.0f) return 2.0f * x;
    if (p == 0.0f) return 0.0f;
    return p * powf(x, p - 1.0f);
}
// d/dp x**p = x**p * log(x)  (autograd pow_backward_exponent; zero where x == 0 and p >= 0)
HODE_HD float dpow_hill_exp(float x, float xp, float p) {
    if (x == 0.0f && p >= 0.0f) return 0.0f;
    return xp * logf(x);
}

// ------------------------------------------------------------------------------------------------------------
// dose schedules (set_action output): one amount + n_dose times per patient
// ------------------------------------------------------------------------------------------------------------
template <int ND>
struct DoseReg {  // times held in registers
    float amt;
    float tau[ND];
    static constexpr int kStatic = ND;
    HODE_HD int n() const { return ND; }
    HODE_HD float at(int j) const { return tau[j]; }
};
struct DoseMem {  // times read from memory (any n_dose)
    float amt;
    const float* tau;
    int nd;
    HODE_HD int n() const { return nd; }
    HODE_HD float at(int j) const { return tau[j]; }
};

// RocheODE.dose_at_time (model.py:509-513): amt * sum_j exp(kel*(tau_j - t) * [t>=tau_j]) * [t>=tau_j]
template <class Dose>
HODE_HD float roche_dose(const Dose& ds, float t, float kel, float kl2) {
    float s = 0.0f;
    const int n = ds.n();
    for (int j = 0; j < n; ++j) {
        const float tau = ds.at(j);
        const float e = exp_scaled(kl2, kel, sub_rn(tau, t));
        s += (t >= tau) ? e : 0.0f;
    }
    return ds.amt * s;
}
// d Dose / d kel = amt * sum_j (tau_j - t) exp(kel (tau_j - t)) [t >= tau_j]
template <class Dose>
HODE_HD float roche_dose_dkel(const Dose& ds, float t, float kel, float kl2) {
    float s = 0.0f;
    const int n = ds.n();
    for (int j = 0; j < n; ++j) {
        const float tau = ds.at(j);
        const float d = sub_rn(tau, t);
        const float e = d * exp_scaled(kl2, kel, d);
        s += (t >= tau) ? e : 0.0f;
    }
    return ds.amt * s;
}
// NeuralODE.dose_at_time (model.py:1015-1017): amt * #{j : tau_j == t}
template <class Dose>
HODE_HD float neural_dose(const Dose& ds, float t) {
    int c = 0;
    const int n = ds.n();
    for (int j = 0; j < n; ++j) c += (ds.at(j) == t) ? 1 : 0;
    return ds.amt * (float)c;
}

// ------------------------------------------------------------------------------------------------------------
// RocheODE: 4 expert states (Disease, ImmuneReact, Immunity, Dose2) + ML latents  (model.py:515-555)
// packed parameters: 13 scalars (named_parameters order), W [ML][D], b [ML]; staged copy appends ec50**HillPatho
// ------------------------------------------------------------------------------------------------------------
enum RocheIdx {
    R_HC = 0, R_HP, R_EC50, R_EMAX, R_KDEXA, R_KDCIR, R_KDCI, R_KDISPROG, R_KID, R_KFB, R_KOFF, R_KIM, R_KEL,
    R_NSCALAR
};

// Warp-cooperative accumulation of the ml_net weight gradients for the wide hybrid field (D = 12: 96 + 8 = 104
// accumulators; kept per thread they push the reverse sweeps to 255 registers plus 1.5 - 1.9 KB of local-memory spills).
// Every lane still owns one trajectory.  Per vjp call each lane publishes u = l (1 - s^2) [ML] and the stage input
// y [D] to a warp-private staging area; after a __syncwarp lane L accumulates, for ITS block of the gradient -- hidden-unit
// pair jb = (L & 7) >> 1, input columns 6 db .. 6 db + 5 with db = L & 1 -- the contributions of 8 of the warp's 32
// trajectories (tg = L >> 3 selects which): 16 LDS.128 + 48 packed FMAs for 14 accumulator registers.  The four partial
// sums per entry meet in shared memory when the kernel flushes.  Call sites must be warp-converged.
template <int D_>
struct RocheCoop {
    static constexpr int ML = D_ - 4;
    static constexpr int NJB = ML / 2;                 // hidden-unit pairs
    static constexpr int SU = 68;                      // floats per unit-pair row: [32 trajectories][2] + pad (bank spread)
    static constexpr int SY = 36;                      // floats per input row: [32 trajectories] + pad
    static constexpr int kStageFloats = NJB * SU + D_ * SY;  // per warp
    static constexpr int NE = 16;                      // expert-scalar accumulators (13 + theta_1, theta_2 of the ablation)
    static_assert(D_ == 12, "lane -> block mapping below is written for D = 12 (4 unit pairs x 2 column halves x 4 groups)");
    float* stage;  // warp-private: U [NJB][32][2] (row stride SU), Y [D][32] (row stride SY)
    int lane;
    float w[6][2];  // dW[2 jb + {0, 1}][6 db + c]
    float b[2];     // db[2 jb + {0, 1}] (identical in the two db lanes; flushed by db == 0)
    float e[NE];    // expert scalars (only touched when EG)
    bool mute;      // true: the call only needs J^T l (zero-weight stage of the continuous adjoint)
};

// ABLATE_: the ablation study's expert part (model.py:545-549): dx = (R, -D theta_1, Q, -I theta_2); theta_1, theta_2 are
// appended to the packed parameters.
template <int D_, bool HILL2_ = false, bool ABLATE_ = false>
struct Roche {
    static constexpr int D = D_;
    static constexpr bool HILL2 = HILL2_;  // caller guarantees HillCure == HillPatho == 2 (checked on the device)
    static constexpr bool ABLATE = ABLATE_;
    static constexpr int ML = D_ - 4;
    static constexpr int OFF_W = R_NSCALAR;
    static constexpr int OFF_B = R_NSCALAR + ML * D_;
    static constexpr int OFF_TH = OFF_B + ML;                 // theta_1, theta_2 (ABLATE only)
    static constexpr int P = OFF_TH + (ABLATE_ ? 2 : 0);      // packed parameter count
    // staged copy appends derived constants, then 8-byte aligned copies of the ml_net weights for the packed (FFMA2)
    // loops: WR = W row-major [ML][D] (pairs over d: the VJP's J^T product), BT = b [ML], WT = W interleaved by unit
    // pairs [ML/2][D][2] (pairs over units: the forward mat-vec keeps each unit's accumulation order)
    static constexpr int OFF_ECP = P;       // ec50 ** HillPatho
    static constexpr int OFF_KL2 = P + 1;   // kel * log2(e)
    static constexpr int OFF_WR = ((P + 2 + 3) / 4) * 4;
    static constexpr int OFF_BT = OFF_WR + ML * D_;
    static constexpr int OFF_WT = OFF_BT + ML;
    static constexpr int SP = OFF_WT + ML * D_;
    static_assert(ML % 2 == 0 && D_ % 2 == 0, "packed loops assume even D and ML");
    static constexpr bool kAccInRegs = true;  // per-thread gradient accumulators fit in registers
    // staged parameters may be read from the constant bank (hode_launch.cuh); the rarely used generic-Hill / ablation variants keep
    // them in shared memory only (half the kernels to build)
    static constexpr bool kConstBank = HILL2_ && !ABLATE_;

    // cooperative copy of one packed parameter set into the staged layout (thread `tid` of `nthr`).
    // With HODE_FOLD_TANH the ml_net weights and biases are pre-multiplied by 2 log2(e): W y + b is then directly the
    // argument of the ex2 inside tanh_pre().
    HODE_HD static void stage(const float* __restrict__ src, float* sp, int tid, int nthr) {
        for (int i = tid; i < P; i += nthr) sp[i] = (i >= OFF_W && i < OFF_TH) ? src[i] * kTanhPre : src[i];
        for (int i = tid; i < ML * D_; i += nthr) {
            const int j = i / D_, d = i % D_;
            const float w = src[OFF_W + i] * kTanhPre;
            sp[OFF_WR + i] = w;
            sp[OFF_WT + ((j >> 1) * D_ + d) * 2 + (j & 1)] = w;
        }
        for (int j = tid; j < ML; j += nthr) sp[OFF_BT + j] = src[OFF_B + j] * kTanhPre;
    }
    HODE_HD static void prepare(float* sp) {
        sp[OFF_ECP] = pow_hill(sp[R_EC50], sp[R_HP]);
        sp[OFF_KL2] = sp[R_KEL] * 1.4426950408889634f;
    }
    // false if this instantiation may not be used with the staged parameters (HILL2 kernels with other exponents)
    template <class PS>
    HODE_HD static bool params_ok(PS sp) { return ABLATE || !HILL2 || (sp[R_HC] == 2.0f && sp[R_HP] == 2.0f); }
    HODE_HD static float hpow(float x, float p) { return HILL2 ? x * x : pow_hill(x, p); }
    HODE_HD static float hdpow(float x, float p) { return HILL2 ? 2.0f * x : dpow_hill(x, p); }

    // f(t, y).  Same terms as model.py:527-544, factored to minimise issue slots (the kernels are issue-bound):
    //   dx1 = D (kp - I^hc kdi - R kdr);  dx2 = D (kid + R kfb) - R (koff + Q kdexa) + R^hp emax / (ec50^hp + R^hp)
    template <class PS, class Dose>
    HODE_HD static void eval(PS sp, float t, const Dose& ds, const float (&y)[D_], float (&dy)[D_]) {
        const float dis = y[0], react = y[1], imm = y[2], dose2 = y[3];
        if (ABLATE) {
            dy[0] = react;
            dy[1] = -dis * sp[OFF_TH];
            dy[2] = dose2;
            dy[3] = -imm * sp[OFF_TH + 1];
        } else {
            const float ip = hpow(imm, sp[R_HC]);
            const float rp = hpow(react, sp[R_HP]);
            dy[0] = dis * fmaf(-react, sp[R_KDCIR], fmaf(-ip, sp[R_KDCI], sp[R_KDISPROG]));
            const float hill = (rp * sp[R_EMAX]) * rcp_f(sp[OFF_ECP] + rp);
            dy[1] = fmaf(dis, fmaf(react, sp[R_KFB], sp[R_KID]), fmaf(-react, fmaf(dose2, sp[R_KDEXA], sp[R_KOFF]), hill));
            dy[2] = react * sp[R_KIM];
            dy[3] = sp[R_KEL] * (roche_dose(ds, t, sp[R_KEL], sp[OFF_KL2]) - dose2);
        }
        // ml_net: two hidden units per packed FMA (weights interleaved by unit pair), same accumulation order per unit
#pragma unroll
        for (int j = 0; j < ML; j += 2) {
            float a0 = sp[OFF_BT + j], a1 = sp[OFF_BT + j + 1];
#pragma unroll
            for (int d = 0; d < D_; ++d)
                fma2s(y[d], sp[OFF_WT + ((j >> 1) * D_ + d) * 2], sp[OFF_WT + ((j >> 1) * D_ + d) * 2 + 1], a0, a1, a0, a1);
            tanh_pre2(a0, a1, dy[4 + j], dy[4 + j + 1]);
        }
    }

    // expert part of gy = J^T l and of the expert-scalar gradients (EG); ML columns of gy are zeroed
    // TH: index of theta_1 in `acc` (OFF_TH in a packed-parameter accumulator, R_NSCALAR in a compact expert-only one)
    template <bool EG, int TH, class PS, class Dose>
    HODE_HD static void vjp_expert(PS sp, float t, const Dose& ds, const float (&y)[D_], const float (&l)[D_],
                                   float (&gy)[D_], float* acc) {
        if (ABLATE) {
            gy[0] = -l[1] * sp[OFF_TH];
            gy[1] = l[0];
            gy[2] = -l[3] * sp[OFF_TH + 1];
            gy[3] = l[2];
#pragma unroll
            for (int d = 4; d < D_; ++d) gy[d] = 0.0f;
            if (EG) {
                acc[TH] -= l[1] * y[0];
                acc[TH + 1] -= l[3] * y[2];
            }
        } else {
            const float dis = y[0], react = y[1], imm = y[2], dose2 = y[3];
            const float hc = sp[R_HC], hp = sp[R_HP], em = sp[R_EMAX], ecp = sp[OFF_ECP];
            const float kdci = sp[R_KDCI], kdcir = sp[R_KDCIR], kfb = sp[R_KFB], kdexa = sp[R_KDEXA];
            const float kel = sp[R_KEL];
            const float ip = hpow(imm, hc);
            const float rp = hpow(react, hp);
            const float inv_den = rcp_f(ecp + rp);
            const float l0 = l[0], l1 = l[1], l2 = l[2], l3 = l[3];
            const float l0d = l0 * dis;
            const float hill_d = (em * ecp) * hdpow(react, hp) * (inv_den * inv_den);  // d/dR of the Hill term
            gy[0] = fmaf(l0, fmaf(-react, kdcir, fmaf(-ip, kdci, sp[R_KDISPROG])), l1 * fmaf(react, kfb, sp[R_KID]));
            gy[1] = fmaf(-l0d, kdcir, fmaf(l1, fmaf(dis, kfb, hill_d) - fmaf(dose2, kdexa, sp[R_KOFF]), l2 * sp[R_KIM]));
            gy[2] = -l0d * kdci * hdpow(imm, hc);
            gy[3] = fmaf(-l1 * react, kdexa, -l3 * kel);
#pragma unroll
            for (int d = 4; d < D_; ++d) gy[d] = 0.0f;
            if (EG) {
                const float ec50 = sp[R_EC50];
                acc[R_KDISPROG] += l0d;
                acc[R_KDCI] -= l0d * ip;
                acc[R_KDCIR] -= l0d * react;
                acc[R_HC] -= l0d * kdci * dpow_hill_exp(imm, ip, hc);
                acc[R_KID] += l1 * dis;
                acc[R_KOFF] -= l1 * react;
                acc[R_KFB] += l1 * dis * react;
                acc[R_EMAX] += l1 * rp * inv_den;
                // d/d ec50 and d/d hp of  em*rp/(ec50**hp + rp)
                acc[R_EC50] -= l1 * em * rp * dpow_hill(ec50, hp) * inv_den * inv_den;
                acc[R_HP] += l1 * em * (dpow_hill_exp(react, rp, hp) * ecp - rp * dpow_hill_exp(ec50, ecp, hp)) * inv_den *
                             inv_den;
                acc[R_KDEXA] -= l1 * dose2 * react;
                acc[R_KIM] += l2 * react;
                const float dose = roche_dose(ds, t, kel, sp[OFF_KL2]);
                acc[R_KEL] += l3 * (dose - dose2 + kel * roche_dose_dkel(ds, t, kel, sp[OFF_KL2]));
            }
        }
    }
    // hidden-unit pair (j, j+1): s = tanh(W y + b) (from k when the caller has f(t, y)), u = l (1 - s^2), gy += W^T u
    template <class PS>
    HODE_HD static void vjp_unit_pair(PS sp, int j, const float (&y)[D_], const float* k, const float (&l)[D_],
                                      float (&gy)[D_], float& u0, float& u1) {
        float s0, s1;
        if (k != nullptr) {
            s0 = k[4 + j]; s1 = k[4 + j + 1];
        } else {
            float a0 = sp[OFF_BT + j], a1 = sp[OFF_BT + j + 1];
#pragma unroll
            for (int d = 0; d < D_; ++d)
                fma2s(y[d], sp[OFF_WT + ((j >> 1) * D_ + d) * 2], sp[OFF_WT + ((j >> 1) * D_ + d) * 2 + 1], a0, a1, a0, a1);
            tanh_pre2(a0, a1, s0, s1);
        }
        // u = l (1 - s^2);  uw = u / kTanhPre (the staged weights carry the factor kTanhPre)
        float q0, q1, uw0, uw1;
        fma2(-s0, -s1, s0, s1, 1.0f, 1.0f, q0, q1);
        mul2(l[4 + j], l[4 + j + 1], q0, q1, u0, u1);
        mul2(u0, u1, kTanhPreInv, kTanhPreInv, uw0, uw1);
#pragma unroll
        for (int d = 0; d < D_; d += 2)
            fma2s(uw0, sp[OFF_WR + j * D_ + d], sp[OFF_WR + j * D_ + d + 1], gy[d], gy[d + 1], gy[d], gy[d + 1]);
#pragma unroll
        for (int d = 0; d < D_; d += 2)
            fma2s(uw1, sp[OFF_WR + (j + 1) * D_ + d], sp[OFF_WR + (j + 1) * D_ + d + 1], gy[d], gy[d + 1], gy[d], gy[d + 1]);
    }

    // gy = J^T l ; acc += d<l, f>/dtheta.  `k` (may be null) is f(t, y) if the caller already has it: its ML part is
    // tanh(W y + b), which is all the MLP backward needs.  EG: also accumulate the 13 expert scalars.
    template <bool EG, class PS, class Dose>
    HODE_HD static void vjp(PS sp, float t, const Dose& ds, const float (&y)[D_], const float* k,
                            const float (&l)[D_], float (&gy)[D_], float* acc) {
        vjp_expert<EG, OFF_TH>(sp, t, ds, y, l, gy, acc);
#pragma unroll
        for (int j = 0; j < ML; j += 2) {
            float u0, u1;
            vjp_unit_pair(sp, j, y, k, l, gy, u0, u1);
#pragma unroll
            for (int d = 0; d < D_; d += 2) {
                fma2s(u0, y[d], y[d + 1], acc[OFF_W + j * D_ + d], acc[OFF_W + j * D_ + d + 1],
                      acc[OFF_W + j * D_ + d], acc[OFF_W + j * D_ + d + 1]);
                fma2s(u1, y[d], y[d + 1], acc[OFF_W + (j + 1) * D_ + d], acc[OFF_W + (j + 1) * D_ + d + 1],
                      acc[OFF_W + (j + 1) * D_ + d], acc[OFF_W + (j + 1) * D_ + d + 1]);
            }
            add2(acc[OFF_B + j], acc[OFF_B + j + 1], u0, u1, acc[OFF_B + j], acc[OFF_B + j + 1]);
        }
    }

#if HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
    // same VJP, ml_net gradients accumulated warp-cooperatively (RocheCoop above); all 32 lanes must call together
    template <bool EG, class PS, class Dose>
    HODE_D static void vjp(PS sp, float t, const Dose& ds, const float (&y)[D_], const float* k, const float (&l)[D_],
                           float (&gy)[D_], RocheCoop<D_>* cp) {
        using C = RocheCoop<D_>;
        vjp_expert<EG, R_NSCALAR>(sp, t, ds, y, l, gy, cp->e);
        float u[ML];
#pragma unroll
        for (int j = 0; j < ML; j += 2) vjp_unit_pair(sp, j, y, k, l, gy, u[j], u[j + 1]);
        if (cp->mute) return;  // warp-uniform
        float* SUp = cp->stage;
        float* SYp = SUp + C::NJB * C::SU;
        const int lane = cp->lane;
        __syncwarp();  // the previous call's owner phase has finished reading the staging area
#pragma unroll
        for (int j = 0; j < ML; j += 2) *reinterpret_cast<float2*>(SUp + (j >> 1) * C::SU + 2 * lane) = make_float2(u[j], u[j + 1]);
#pragma unroll
        for (int d = 0; d < D_; ++d) SYp[d * C::SY + lane] = y[d];
        __syncwarp();
        // owner phase: unit pair jb, columns 6 db .. 6 db + 5, trajectories 4 tg + 16 h + {0..3}
        const int tg = lane >> 3, jb = (lane & 7) >> 1, db = lane & 1;
        const float* up = SUp + jb * C::SU + 8 * tg;
        const float* yp = SYp + (6 * db) * C::SY + 4 * tg;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 ua = *reinterpret_cast<const float4*>(up + 32 * h);      // (u_j0, u_j1) of trajectories tt, tt+1
            const float4 ub = *reinterpret_cast<const float4*>(up + 32 * h + 4);  // ... tt+2, tt+3
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const float4 yv = *reinterpret_cast<const float4*>(yp + c * C::SY + 16 * h);
                fma2s(yv.x, ua.x, ua.y, cp->w[c][0], cp->w[c][1], cp->w[c][0], cp->w[c][1]);
                fma2s(yv.y, ua.z, ua.w, cp->w[c][0], cp->w[c][1], cp->w[c][0], cp->w[c][1]);
                fma2s(yv.z, ub.x, ub.y, cp->w[c][0], cp->w[c][1], cp->w[c][0], cp->w[c][1]);
                fma2s(yv.w, ub.z, ub.w, cp->w[c][0], cp->w[c][1], cp->w[c][0], cp->w[c][1]);
            }
            add2(cp->b[0], cp->b[1], ua.x, ua.y, cp->b[0], cp->b[1]);
            add2(cp->b[0], cp->b[1], ua.z, ua.w, cp->b[0], cp->b[1]);
            add2(cp->b[0], cp->b[1], ub.x, ub.y, cp->b[0], cp->b[1]);
            add2(cp->b[0], cp->b[1], ub.z, ub.w, cp->b[0], cp->b[1]);
        }
    }
#endif
};

// ------------------------------------------------------------------------------------------------------------
// NeuralODE: dy = tanh(W2 tanh(W1 [y, Dose] + b1) + b2), hidden width 10*D  (model.py:969-1026)
// packed parameters: kel (unused by the field but a state_dict key), W1 [H][D+1], b1 [H], W2 [D][H], b2 [D].
// staged layout: one 16-byte aligned record per hidden unit j: {W1[j][0..D], b1[j], W2[0..D-1][j], pad}, then b2 --
// every weight the j-th unit needs is contiguous, so the broadcast reads are LDS.128.
// ------------------------------------------------------------------------------------------------------------
// Warp-cooperative parameter-gradient accumulation for the NeuralODE field (846 - 3 132 parameters: per-thread
// accumulators would live in local memory and make the reverse sweep ~35x slower than the forward sweep).
// Every lane still integrates its own trajectory, but the weight gradients are owned by HIDDEN UNIT: lane L keeps the
// rows of dW1 / db1 / dW2 of the units j = 32 c + L in registers.  Per vjp call and per chunk of 32 units every lane
// publishes its (delta_j, a_j) -- and once per call its (input, output-adjoint) vectors -- to a warp-private staging area
// in shared memory; after a __syncwarp each lane walks the 32 trajectories of the warp and accumulates the outer
// products of ITS unit.  The call sites must be warp-converged (fixed-grid sweeps, batch-coupled dopri5).
template <int D_>
struct NeuralCoop {
    static constexpr int H = 10 * D_, IN = D_ + 1, NJ = (H + 31) / 32;
    static constexpr int SI = (D_ + 2 + 3) / 4 * 4;  // staged input row: in[0..D], 1 (the bias "input"), zero padding
    static constexpr int SU = (D_ + 3) / 4 * 4;      // staged output-adjoint row
    static constexpr int SDA = 66;                   // (delta, a) pairs of one unit over the 32 lanes, +2 floats: conflict-free LDS.64
    static constexpr int kStageFloats = (SI + SU) * 32 + 32 * SDA;  // per warp
    float* stage;          // warp-private: S_in [32][SI], S_u [32][SU], S_da [32 units][SDA]
    int lane;
    float w1[NJ][SI];      // rows of dW1 (IN) and db1 (1) of the owned units (+ padding that stays zero)
    float w2[NJ][SU];      // columns of dW2 of the owned units
    float b2[D_];          // this lane's own contribution to db2 (reduced over the warp at the end)
    bool mute;             // true: the call only needs J^T l (zero-weight stage of the continuous adjoint)
};

// The two mat-vecs are written on packed pairs (FFMA2, see above): layer 1 as a dot product over input pairs
// (w1[2q], w1[2q+1]) . (in[2q], in[2q+1]) with the bias riding along as the weight of a constant-one input, layer 2 as
// (out[2q], out[2q+1]) += a_j (W2[2q][j], W2[2q+1][j]).  13 -> 8 math issue slots per hidden unit at D = 6.  Both layers'
// staged weights carry the factor 2 log2(e) of the tanh (HODE_FOLD_TANH), which the VJP folds back into its cotangents.
template <int D_>
struct Neural {
    static constexpr int D = D_;
    static constexpr int H = 10 * D_;
    static constexpr int IN = D_ + 1;
    static constexpr int P = 1 + H * IN + H + D_ * H + D_;
    static constexpr int NP1 = (D_ + 2) / 2;              // input pairs of layer 1: in[0..D], 1
    static constexpr int R = ((2 * D_ + 2 + 3) / 4) * 4;  // record length
    static constexpr int SP = H * R + ((D_ + 3) / 4) * 4;
    static constexpr int OFF_W1 = 1;
    static constexpr int OFF_B1 = 1 + H * IN;
    static constexpr int OFF_W2 = OFF_B1 + H;
    static constexpr int OFF_B2 = OFF_W2 + D_ * H;
    static constexpr bool kAccInRegs = false;  // P is 846..3132: accumulators live in local memory
    static constexpr bool kConstBank = false;
    static_assert(D_ % 2 == 0, "packed pairs assume an even state dimension");

    // staged record of hidden unit j: {W1[j][0..D], b1[j], W2[0..D-1][j], pad} * kTanhPre, then b2 * kTanhPre
    HODE_HD static void stage(const float* __restrict__ src, float* sp, int tid, int nthr) {
        for (int e = tid; e < H * R; e += nthr) {
            const int j = e / R, c = e % R;
            float v = 0.0f;
            if (c < IN) v = src[OFF_W1 + j * IN + c];
            else if (c == IN) v = src[OFF_B1 + j];
            else if (c < IN + 1 + D_) v = src[OFF_W2 + (c - IN - 1) * H + j];
            sp[e] = v * kTanhPre;
        }
        for (int d = tid; d < D_; d += nthr) sp[H * R + d] = src[OFF_B2 + d] * kTanhPre;
    }
    HODE_HD static void prepare(float*) {}

    template <class PS>
    HODE_HD static bool params_ok(PS) { return true; }

    HODE_HD static float act(float pre) {  // tanh of a pre-activation that already carries kTanhPre
#if HODE_FOLD_TANH
        return tanh_pre(pre);
#else
        return tanh_f(pre);
#endif
    }
    // (in[0..D], 1): the layer-1 input with the bias input appended
    template <class Dose>
    HODE_HD static void inputs(float t, const Dose& ds, const float (&y)[D_], float (&in)[2 * NP1]) {
#pragma unroll
        for (int i = 0; i < D_; ++i) in[i] = y[i];
        in[D_] = neural_dose(ds, t);
        in[D_ + 1] = 1.0f;
    }
    // hidden activation of one unit from its record
    HODE_HD static float hidden(const float* rec, const float (&in)[2 * NP1]) {
        float p0 = 0.0f, p1 = 0.0f;
#pragma unroll
        for (int q = 0; q < NP1; ++q) fma2(rec[2 * q], rec[2 * q + 1], in[2 * q], in[2 * q + 1], p0, p1, p0, p1);
        return act(p0 + p1);
    }

    template <class Dose>
    HODE_HD static void eval(const float* __restrict__ sp, float t, const Dose& ds, const float (&y)[D_],
                             float (&dy)[D_]) {
        float in[2 * NP1], out[D_];
        inputs(t, ds, y, in);
#pragma unroll
        for (int d = 0; d < D_; ++d) out[d] = sp[H * R + d];
#pragma unroll 4
        for (int j = 0; j < H; ++j) {
            const float* rec = sp + j * R;
            const float a = hidden(rec, in);
#pragma unroll
            for (int d = 0; d < D_; d += 2) fma2s(a, rec[IN + 1 + d], rec[IN + 2 + d], out[d], out[d + 1], out[d], out[d + 1]);
        }
#pragma unroll
        for (int d = 0; d < D_; ++d) dy[d] = act(out[d]);
    }

    // u = l (1 - s^2): cotangent of the output pre-activation; `us` additionally carries kTanhPre so that it can be
    // contracted with the STAGED second-layer weights
    template <class Dose>
    HODE_HD static void out_cotangent(const float* __restrict__ sp, float t, const Dose& ds, const float (&y)[D_],
                                      const float* k, const float (&l)[D_], float (&u)[D_]) {
        if (k != nullptr) {
#pragma unroll
            for (int d = 0; d < D_; ++d) u[d] = l[d] * (1.0f - k[d] * k[d]);
        } else {
            float s2[D_];
            eval(sp, t, ds, y, s2);
#pragma unroll
            for (int d = 0; d < D_; ++d) u[d] = l[d] * (1.0f - s2[d] * s2[d]);
        }
    }
    // one hidden unit of the VJP: activation a, delta = (W2[:, j] . u) (1 - a^2), gy += W1[j][0..D-1] delta
    HODE_HD static void unit_vjp(const float* rec, const float (&in)[2 * NP1], const float (&us)[D_], float (&gy)[D_],
                                 float& a, float& del) {
        a = hidden(rec, in);
        float c0 = 0.0f, c1 = 0.0f;
#pragma unroll
        for (int d = 0; d < D_; d += 2) fma2(rec[IN + 1 + d], rec[IN + 2 + d], us[d], us[d + 1], c0, c1, c0, c1);
        del = (c0 + c1) * (1.0f - a * a);
        // gy is accumulated against the STAGED first-layer weights (W1 * kTanhPre); the callers rescale it once at the end
#pragma unroll
        for (int i = 0; i < D_; i += 2) fma2s(del, rec[i], rec[i + 1], gy[i], gy[i + 1], gy[i], gy[i + 1]);
    }

    template <bool EG, class Dose>
    HODE_HD static void vjp(const float* __restrict__ sp, float t, const Dose& ds, const float (&y)[D_],
                            const float* k, const float (&l)[D_], float (&gy)[D_], float* acc) {
        float in[2 * NP1], u[D_], us[D_];
        inputs(t, ds, y, in);
        out_cotangent(sp, t, ds, y, k, l, u);
#pragma unroll
        for (int d = 0; d < D_; ++d) { acc[OFF_B2 + d] += u[d]; gy[d] = 0.0f; us[d] = u[d] * kTanhPreInv; }
#pragma unroll 1
        for (int j = 0; j < H; ++j) {
            const float* rec = sp + j * R;
            float a, del;
            unit_vjp(rec, in, us, gy, a, del);
#pragma unroll
            for (int d = 0; d < D_; ++d) acc[OFF_W2 + d * H + j] = fmaf(u[d], a, acc[OFF_W2 + d * H + j]);
#pragma unroll
            for (int i = 0; i < IN; ++i) acc[OFF_W1 + j * IN + i] = fmaf(del, in[i], acc[OFF_W1 + j * IN + i]);
            acc[OFF_B1 + j] += del;
        }
#pragma unroll
        for (int i = 0; i < D_; ++i) gy[i] *= kTanhPreInv;
    }

#if HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
    // same VJP, parameter gradients accumulated warp-cooperatively (NeuralCoop above); all 32 lanes must call together
    template <bool EG, class Dose>
    HODE_D static void vjp(const float* __restrict__ sp, float t, const Dose& ds, const float (&y)[D_], const float* k,
                           const float (&l)[D_], float (&gy)[D_], NeuralCoop<D_>* cp) {
        using C = NeuralCoop<D_>;
        float in[2 * NP1], u[D_], us[D_];
        inputs(t, ds, y, in);
        out_cotangent(sp, t, ds, y, k, l, u);
        float* S_in = cp->stage;
        float* S_u = S_in + C::SI * 32;
        float* S_da = S_u + C::SU * 32;
        const int lane = cp->lane;
        __syncwarp();  // the previous call's owner phase has finished reading the staging area
        // this lane's rows of the staging area: the layer-1 input (with the constant-one bias input) and u, zero padded
        {
            float row[C::SI];
#pragma unroll
            for (int i = 0; i < C::SI; ++i) row[i] = i < 2 * NP1 ? in[i] : 0.0f;
#pragma unroll
            for (int i = 0; i < C::SI; i += 4) *reinterpret_cast<float4*>(S_in + lane * C::SI + i) = make_float4(row[i], row[i + 1], row[i + 2], row[i + 3]);
            float ru[C::SU];
#pragma unroll
            for (int d = 0; d < C::SU; ++d) ru[d] = d < D_ ? u[d] : 0.0f;
#pragma unroll
            for (int d = 0; d < C::SU; d += 4) *reinterpret_cast<float4*>(S_u + lane * C::SU + d) = make_float4(ru[d], ru[d + 1], ru[d + 2], ru[d + 3]);
        }
#pragma unroll
        for (int d = 0; d < D_; ++d) { cp->b2[d] += cp->mute ? 0.0f : u[d]; gy[d] = 0.0f; us[d] = u[d] * kTanhPreInv; }
#pragma unroll
        for (int c = 0; c < C::NJ; ++c) {
            // producer phase: this lane's trajectory, hidden units 32 c .. 32 c + 31
#pragma unroll 1
            for (int jj = 0; jj < 32; ++jj) {
                const int j = c * 32 + jj;
                float del = 0.0f, a = 0.0f;
                if (j < H) unit_vjp(sp + j * R, in, us, gy, a, del);
                *reinterpret_cast<float2*>(S_da + jj * C::SDA + 2 * lane) = make_float2(del, a);
            }
            __syncwarp();
            if (cp->mute) { __syncwarp(); continue; }  // warp-uniform
            // owner phase: unit j = 32 c + lane, all 32 trajectories of the warp: dW1[j][:] += delta in, db1[j] += delta (the
            // constant-one input), dW2[:, j] += a u -- vector loads of the staged rows (broadcast), packed FMAs
#pragma unroll 4
            for (int tt = 0; tt < 32; ++tt) {
                const float2 da = *reinterpret_cast<const float2*>(S_da + lane * C::SDA + 2 * tt);
#pragma unroll
                for (int i = 0; i < C::SI; i += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(S_in + tt * C::SI + i);
                    fma2s(da.x, v.x, v.y, cp->w1[c][i], cp->w1[c][i + 1], cp->w1[c][i], cp->w1[c][i + 1]);
                    fma2s(da.x, v.z, v.w, cp->w1[c][i + 2], cp->w1[c][i + 3], cp->w1[c][i + 2], cp->w1[c][i + 3]);
                }
#pragma unroll
                for (int d = 0; d < C::SU; d += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(S_u + tt * C::SU + d);
                    fma2s(da.y, v.x, v.y, cp->w2[c][d], cp->w2[c][d + 1], cp->w2[c][d], cp->w2[c][d + 1]);
                    fma2s(da.y, v.z, v.w, cp->w2[c][d + 2], cp->w2[c][d + 3], cp->w2[c][d + 2], cp->w2[c][d + 3]);
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < D_; ++i) gy[i] *= kTanhPreInv;
    }
#endif
};

// ------------------------------------------------------------------------------------------------------------
// fixed-grid step functions (tde fixed_grid.py): dy such that y1 = y0 + dy.  Time arithmetic in float32.
// ------------------------------------------------------------------------------------------------------------
enum Method { M_EULER = 0, M_MIDPOINT = 1, M_RK4_38 = 2, M_DOPRI5 = 3 };

#define HODE_ONE_THIRD ((float)(1.0 / 3.0))
#define HODE_TWO_THIRDS ((float)(2.0 / 3.0))

template <class F, int METHOD, class PS, class Dose>
HODE_HD void fixed_step(PS sp, const Dose& ds, float t0, float t1, float dt, bool perturb,
                        const float (&y0)[F::D], float (&y1)[F::D]) {
    constexpr int D = F::D;
    float k1[D];
    F::eval(sp, perturb ? t_next(t0) : t0, ds, y0, k1);
    if (METHOD == M_EULER) {
#pragma unroll
        for (int d = 0; d < D; ++d) y1[d] = y0[d] + dt * k1[d];
    } else if (METHOD == M_MIDPOINT) {
        const float half_dt = mul_rn(0.5f, dt);
        float ym[D], k2[D];
#pragma unroll
        for (int d = 0; d < D; ++d) ym[d] = y0[d] + k1[d] * half_dt;
        F::eval(sp, add_rn(t0, half_dt), ds, ym, k2);
#pragma unroll
        for (int d = 0; d < D; ++d) y1[d] = y0[d] + dt * k2[d];
    } else {  // 3/8 rule (tde rk4_alt_step_func), written with fused multiply-adds
        const float c13 = dt * HODE_ONE_THIRD, w1 = dt * 0.125f, w3 = dt * 0.375f;
        float yi[D], k2[D], k3[D], k4[D], tmp[D];
        v_axpy<D>(yi, c13, k1, y0);
        F::eval(sp, add_rn(t0, mul_rn(dt, HODE_ONE_THIRD)), ds, yi, k2);
        v_axpy<D>(tmp, -c13, k1, y0);
        v_axpy<D>(yi, dt, k2, tmp);
        F::eval(sp, add_rn(t0, mul_rn(dt, HODE_TWO_THIRDS)), ds, yi, k3);
        v_sub<D>(tmp, k1, k2);
        v_add<D>(tmp, tmp, k3);
        v_axpy<D>(yi, dt, tmp, y0);
        F::eval(sp, perturb ? t_prev(t1) : t1, ds, yi, k4);
        v_add<D>(tmp, k2, k3);
        v_axpy<D>(tmp, w3, tmp, y0);
        v_add<D>(yi, k1, k4);
        v_axpy<D>(y1, w1, yi, tmp);
    }
}

// reverse of one fixed-grid step:  lam0 = lam1 + (d dy/d y0)^T lam1 ; acc += (d dy/d theta)^T lam1.
// Stages are recomputed from y0 (the tape holds only the state at the start of the step).
template <class F, int METHOD, bool EG, class PS, class Dose, class ACC>
HODE_HD void fixed_step_vjp(PS sp, const Dose& ds, float t0, float t1, float dt, bool perturb,
                            const float (&y0)[F::D], const float (&lam1)[F::D], float (&lam0)[F::D], ACC acc) {
    constexpr int D = F::D;
    const float ta = perturb ? t_next(t0) : t0;
    float kb[D], g[D];
    if (METHOD == M_EULER) {
#pragma unroll
        for (int d = 0; d < D; ++d) kb[d] = dt * lam1[d];
        F::template vjp<EG>(sp, ta, ds, y0, nullptr, kb, g, acc);
#pragma unroll
        for (int d = 0; d < D; ++d) lam0[d] = lam1[d] + g[d];
    } else if (METHOD == M_MIDPOINT) {
        const float half_dt = mul_rn(0.5f, dt);
        float k1[D], ym[D];
        F::eval(sp, ta, ds, y0, k1);
#pragma unroll
        for (int d = 0; d < D; ++d) ym[d] = y0[d] + k1[d] * half_dt;
#pragma unroll
        for (int d = 0; d < D; ++d) kb[d] = dt * lam1[d];
        F::template vjp<EG>(sp, add_rn(t0, half_dt), ds, ym, nullptr, kb, g, acc);  // g = adjoint of ym
#pragma unroll
        for (int d = 0; d < D; ++d) {
            lam0[d] = lam1[d] + g[d];
            kb[d] = half_dt * g[d];  // adjoint of k1
        }
        F::template vjp<EG>(sp, ta, ds, y0, k1, kb, g, acc);
#pragma unroll
        for (int d = 0; d < D; ++d) lam0[d] += g[d];
    } else {
        // Butcher form of the 3/8 rule: A = [[],[1/3],[-1/3,1],[1,-1,1]], b = [1/8,3/8,3/8,1/8].
        // Stage inputs are rebuilt from k1..k3 when needed instead of being kept (register pressure).
        const float tb = add_rn(t0, mul_rn(dt, HODE_ONE_THIRD));
        const float tc = add_rn(t0, mul_rn(dt, HODE_TWO_THIRDS));
        const float td = perturb ? t_prev(t1) : t1;
        const float c13 = dt * HODE_ONE_THIRD, w1 = dt * 0.125f, w3 = dt * 0.375f;
        float k1[D], k2[D], k3[D], Y[D], g4[D], g3[D], tmp[D];
        F::eval(sp, ta, ds, y0, k1);
        v_axpy<D>(Y, c13, k1, y0);
        F::eval(sp, tb, ds, Y, k2);
        v_axpy<D>(tmp, -c13, k1, y0);
        v_axpy<D>(Y, dt, k2, tmp);
        F::eval(sp, tc, ds, Y, k3);
        // stage 4: kb4 = b4 dt lam1
        v_sub<D>(tmp, k1, k2);
        v_add<D>(tmp, tmp, k3);
        v_axpy<D>(Y, dt, tmp, y0);
        v_scale<D>(kb, w1, lam1);
        F::template vjp<EG>(sp, td, ds, Y, nullptr, kb, g4, acc);
        // stage 3: kb3 = b3 dt lam1 + a43 dt g4
        v_add<D>(lam0, lam1, g4);
        v_axpy<D>(tmp, -c13, k1, y0);
        v_axpy<D>(Y, dt, k2, tmp);
        v_scale<D>(tmp, w3, lam1);
        v_axpy<D>(kb, dt, g4, tmp);
        F::template vjp<EG>(sp, tc, ds, Y, k3, kb, g3, acc);
        // stage 2: kb2 = b2 dt lam1 + a42 dt g4 + a32 dt g3 = 3/8 dt lam1 + dt (g3 - g4)
        v_add<D>(lam0, lam0, g3);
        v_axpy<D>(Y, c13, k1, y0);
        v_sub<D>(kb, g3, g4);
        v_axpy<D>(kb, dt, kb, tmp);  // tmp still holds w3 lam1
        F::template vjp<EG>(sp, tb, ds, Y, k2, kb, g, acc);
        // stage 1: kb1 = b1 dt lam1 + a41 dt g4 + a31 dt g3 + a21 dt g2 = 1/8 dt lam1 + dt g4 + dt/3 (g2 - g3)
        v_add<D>(lam0, lam0, g);
        v_scale<D>(tmp, w1, lam1);
        v_axpy<D>(tmp, dt, g4, tmp);
        v_sub<D>(kb, g, g3);
        v_axpy<D>(kb, c13, kb, tmp);
        F::template vjp<EG>(sp, ta, ds, y0, k1, kb, g, acc);
        v_add<D>(lam0, lam0, g);
    }
}

// ------------------------------------------------------------------------------------------------------------
// continuous adjoint (tde adjoint.py OdeintAdjointMethod.backward): the augmented system (y, a, g_theta) is integrated
// BACKWARDS over every output interval with the same fixed-grid method.  tde reverses time by negation
// (odeint.py _ReverseFunc): in s = -t the system is  dy/ds = -f(-s, y),  da/ds = +J^T a,  dg/ds = +(df/dtheta)^T a.
// ------------------------------------------------------------------------------------------------------------
// J^T l only (a stage whose quadrature weight is zero): the parameter part goes to a dead local array / a muted
// cooperative accumulator
template <class F, class PS, class Dose>
HODE_HD void vjp_state_only(PS sp, float t, const Dose& ds, const float (&y)[F::D], const float* k,
                            const float (&l)[F::D], float (&gy)[F::D], float* /*acc*/) {
    float dead[F::P];
#pragma unroll
    for (int p = 0; p < F::P; ++p) dead[p] = 0.0f;
    F::template vjp<false>(sp, t, ds, y, k, l, gy, (float*)dead);
}
#if HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
template <class F, class PS, class Dose, int DD>
HODE_D void vjp_state_only(PS sp, float t, const Dose& ds, const float (&y)[F::D], const float* k,
                           const float (&l)[F::D], float (&gy)[F::D], NeuralCoop<DD>* cp) {
    cp->mute = true;
    F::template vjp<false>(sp, t, ds, y, k, l, gy, cp);
    cp->mute = false;
}
template <class F, class PS, class Dose, int DD>
HODE_D void vjp_state_only(PS sp, float t, const Dose& ds, const float (&y)[F::D], const float* k,
                           const float (&l)[F::D], float (&gy)[F::D], RocheCoop<DD>* cp) {
    cp->mute = true;
    F::template vjp<false>(sp, t, ds, y, k, l, gy, cp);
    cp->mute = false;
}
#endif

// One step s0 -> s1 (ds = s1 - s0 > 0) of the augmented system in negated time; y and a are updated in place and
// acc += ds * sum_i b_i (df/dtheta(y_i))^T a_i.  The vjp is linear in its cotangent, so stage i is called with
// l_i = (b_i ds) a_i: the accumulators then receive the quadrature-weighted term directly and the returned
// g_i = (b_i ds) J_i^T a_i enters the next stage adjoints with dt-free constant ratios (a_ij / b_j).
template <class F, int METHOD, bool EG, class PS, class Dose, class ACC>
HODE_HD void fixed_adjoint_step(PS sp, const Dose& ds, float s0, float s1, float dt, bool perturb,
                                float (&y)[F::D], float (&a)[F::D], ACC acc) {
    constexpr int D = F::D;
    const float ta = -(perturb ? t_next(s0) : s0);
    float k1[D], l[D], g1[D];
    F::eval(sp, ta, ds, y, k1);
    if (METHOD == M_EULER) {
#pragma unroll
        for (int d = 0; d < D; ++d) l[d] = dt * a[d];
        F::template vjp<EG>(sp, ta, ds, y, k1, l, g1, acc);
#pragma unroll
        for (int d = 0; d < D; ++d) { y[d] = fmaf(-dt, k1[d], y[d]); a[d] += g1[d]; }
    } else if (METHOD == M_MIDPOINT) {
        const float half_dt = mul_rn(0.5f, dt);
        const float tm = -add_rn(s0, half_dt);
        float ym[D], am[D], k2[D], g2[D];
#pragma unroll
        for (int d = 0; d < D; ++d) l[d] = half_dt * a[d];
        vjp_state_only<F>(sp, ta, ds, y, k1, l, g1, acc);  // b_1 = 0
#pragma unroll
        for (int d = 0; d < D; ++d) { ym[d] = fmaf(-half_dt, k1[d], y[d]); am[d] = a[d] + g1[d]; }
        F::eval(sp, tm, ds, ym, k2);
#pragma unroll
        for (int d = 0; d < D; ++d) l[d] = dt * am[d];
        F::template vjp<EG>(sp, tm, ds, ym, k2, l, g2, acc);
#pragma unroll
        for (int d = 0; d < D; ++d) { y[d] = fmaf(-dt, k2[d], y[d]); a[d] += g2[d]; }
    } else {  // 3/8 rule: A = [[],[1/3],[-1/3,1],[1,-1,1]], b = [1/8,3/8,3/8,1/8]
        const float tb = -add_rn(s0, mul_rn(dt, HODE_ONE_THIRD));
        const float tc = -add_rn(s0, mul_rn(dt, HODE_TWO_THIRDS));
        const float td = -(perturb ? t_prev(s1) : s1);
        const float c13 = dt * HODE_ONE_THIRD, w1 = dt * 0.125f, w3 = dt * 0.375f;
        constexpr float r83 = (float)(8.0 / 3.0);
        // running sums keep the live set small (D = 12 carries 117 gradient accumulators next to the state):
        //   ys = y - ds sum_i b_i k_i,  as = a + sum_i g_i,  ks = k1 - k2 (+ k3),  ga = a + 8 g1 - 8/3 g2
        float Y[D], ys[D], as[D], ks[D], ga[D], kk[D], g[D];
        v_scale<D>(l, w1, a);
        F::template vjp<EG>(sp, ta, ds, y, k1, l, g1, acc);
        v_axpy<D>(ys, -w1, k1, y);
        v_add<D>(as, a, g1);
        // stage 2: y2 = y - ds/3 k1 ; a2 = a + ds/3 J1^T a = a + 8/3 g1
        v_axpy<D>(Y, -c13, k1, y);
        v_axpy<D>(ga, r83, g1, a);
        v_scale<D>(l, w3, ga);
        F::eval(sp, tb, ds, Y, kk);
        F::template vjp<EG>(sp, tb, ds, Y, kk, l, g, acc);
        v_axpy<D>(ys, -w3, kk, ys);
        v_add<D>(as, as, g);
        // stage 3: y3 = y - ds (k2 - k1/3) ; a3 = a + 8/3 (g2 - g1)
        v_axpy<D>(Y, c13, k1, y);
        v_axpy<D>(Y, -dt, kk, Y);
        v_sub<D>(ks, k1, kk);
        v_sub<D>(ga, g, g1);
        v_axpy<D>(ga, r83, ga, a);
        v_scale<D>(l, w3, ga);
        v_axpy<D>(ga, 8.0f, g1, a);    // ga = a + 8 g1 - 8/3 g2
        v_axpy<D>(ga, -r83, g, ga);
        F::eval(sp, tc, ds, Y, kk);
        F::template vjp<EG>(sp, tc, ds, Y, kk, l, g, acc);
        v_axpy<D>(ys, -w3, kk, ys);
        v_add<D>(as, as, g);
        // stage 4: y4 = y - ds (k1 - k2 + k3) ; a4 = a + 8 g1 - 8/3 g2 + 8/3 g3
        v_add<D>(ks, ks, kk);
        v_axpy<D>(Y, -dt, ks, y);
        v_axpy<D>(ga, r83, g, ga);
        v_scale<D>(l, w1, ga);
        F::eval(sp, td, ds, Y, kk);
        F::template vjp<EG>(sp, td, ds, Y, kk, l, g, acc);
        v_axpy<D>(y, -w1, kk, ys);
        v_add<D>(a, as, g);
    }
}

// ------------------------------------------------------------------------------------------------------------
// dopri5 (tde dopri5.py tableau, cast to float32 like `tableau.to(y0.dtype)`).  The coefficients live in tables that can
// be indexed at run time: the D = 12 kernels keep their stage derivatives in shared memory and run the stage loops
// ROLLED (one inlined copy of the vector field instead of seven: the unrolled reverse sweep was 118 KB of code, far
// beyond the 32 KB instruction cache -- ncu: stall_no_instruction 2.1 per issue), the small-D kernels unroll them and
// the compiler folds the constants.
// ------------------------------------------------------------------------------------------------------------
#define HODE_D5_ALPHA {(float)(1.0 / 5), (float)(3.0 / 10), (float)(4.0 / 5), (float)(8.0 / 9), 1.0f, 1.0f}
#define HODE_D5_BETA                                                                                                        \
    {{(float)(1.0 / 5), 0, 0, 0, 0, 0},                                                                                     \
     {(float)(3.0 / 40), (float)(9.0 / 40), 0, 0, 0, 0},                                                                    \
     {(float)(44.0 / 45), (float)(-56.0 / 15), (float)(32.0 / 9), 0, 0, 0},                                                 \
     {(float)(19372.0 / 6561), (float)(-25360.0 / 2187), (float)(64448.0 / 6561), (float)(-212.0 / 729), 0, 0},             \
     {(float)(9017.0 / 3168), (float)(-355.0 / 33), (float)(46732.0 / 5247), (float)(49.0 / 176), (float)(-5103.0 / 18656), 0}, \
     {(float)(35.0 / 384), 0.0f, (float)(500.0 / 1113), (float)(125.0 / 192), (float)(-2187.0 / 6784), (float)(11.0 / 84)}}
#define HODE_D5_CERR                                                                                                     \
    {(float)(35.0 / 384 - 1951.0 / 21600), 0.0f, (float)(500.0 / 1113 - 22642.0 / 50085), (float)(125.0 / 192 - 451.0 / 720), \
     (float)(-2187.0 / 6784 - -12231.0 / 42400), (float)(11.0 / 84 - 649.0 / 6300), (float)(-1.0 / 60.0)}
#define HODE_D5_CMID                                                                                                      \
    {(float)(6025192743.0 / 30085553152.0 / 2), 0.0f, (float)(51252292925.0 / 65400821598.0 / 2),                         \
     (float)(-2691868925.0 / 45128329728.0 / 2), (float)(187940372067.0 / 1594534317056.0 / 2),                           \
     (float)(-1776094331.0 / 19743644256.0 / 2), (float)(11237099.0 / 235043384.0 / 2)}
// host copies (the plain-C++ build, and the never-executed host pass of the __host__ __device__ bodies under nvcc)
static const float kD5AlphaH[6] = HODE_D5_ALPHA;
static const float kD5BetaH[6][6] = HODE_D5_BETA;
static const float kD5CErrH[7] = HODE_D5_CERR;
static const float kD5CMidH[7] = HODE_D5_CMID;
#if HODE_DEVICE_BUILD
static __constant__ float kD5Alpha[6] = HODE_D5_ALPHA;
static __constant__ float kD5Beta[6][6] = HODE_D5_BETA;
static __constant__ float kD5CErr[7] = HODE_D5_CERR;
static __constant__ float kD5CMid[7] = HODE_D5_CMID;
#endif
#if HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
HODE_D float d5_alpha(int i) { return kD5Alpha[i]; }
HODE_D float d5_beta(int i, int j) { return kD5Beta[i][j]; }
HODE_D float d5_cerr(int i) { return kD5CErr[i]; }
HODE_D float d5_cmid(int i) { return kD5CMid[i]; }
#else
HODE_HD float d5_alpha(int i) { return kD5AlphaH[i]; }
HODE_HD float d5_beta(int i, int j) { return kD5BetaH[i][j]; }
HODE_HD float d5_cerr(int i) { return kD5CErrH[i]; }
HODE_HD float d5_cmid(int i) { return kD5CMidH[i]; }
#endif

// torch.max / torch.min propagate NaN (fmaxf / fminf drop it)
#if HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
HODE_D float nan_maxf(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
HODE_D float nan_minf(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
#else
HODE_HD float nan_maxf(float a, float b) { return (a != a || b != b) ? (a + b) : (a > b ? a : b); }
HODE_HD float nan_minf(float a, float b) { return (a != a || b != b) ? (a + b) : (a < b ? a : b); }
#endif

// tde misc.py _optimal_step_size (order = 5): dt * min(ifactor, max(safety / ratio**(1/5), dfactor or 1)).  The ratio is a
// float32 quantity (tde computes it in y.dtype); the factor is evaluated in float32 as 2^(-log2(ratio) / 5) with the MUFU
// lg2 / ex2 units (relative error ~3e-7, the effect of a few ulps of the ratio itself) and only the product with dt is
// float64 -- the float64 pow + divide of the literal formula were ~250 instructions per attempt.  ratio == 0 gives
// 2^(+inf) = inf -> min(ifactor, inf) = ifactor like tde's special case; NaN propagates like torch.min / torch.max.
#if HODE_FAST_MATH && HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
HODE_D float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
HODE_D float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
HODE_D float inv_fifth_root(float r) { return ex2_approx(-0.2f * lg2_approx(r)); }
HODE_D float fast_sqrt(float x) { return sqrt_approx(x); }
// e / tol for tol >= atol > 0 (one MUFU.RCP + one multiply instead of the ~12-instruction IEEE division)
HODE_D float div_tol(float e, float tol) { return e * rcp_approx(tol); }
#else
HODE_HD float inv_fifth_root(float r) { return 1.0f / powf(r, 0.2f); }
HODE_HD float fast_sqrt(float x) { return sqrtf(x); }
HODE_HD float div_tol(float e, float tol) { return e / tol; }
#endif
HODE_HD double optimal_step(double last, float ratio, float safety, float ifactor, float dfactor) {
    if (ratio < 1.0f) dfactor = 1.0f;
    const float factor = nan_minf(ifactor, nan_maxf(safety * inv_fifth_root(ratio), dfactor));
    return last * (double)factor;
}

// ------------------------------------------------------------------------------------------------------------
// Per-thread rows of D floats (stage derivatives, stage adjoints).  RowsReg keeps them in registers: every index must
// be a compile-time constant after unrolling (kDynamic = false).  RowsMem keeps them in the thread's slots of a
// shared-memory array, interleaved by thread so that a warp's vector access touches consecutive 16 / 8-byte chunks
// (conflict-free): chunk c of row i of this thread is at p[(i * NC + c) * stride]; rows may be indexed at run time.
// ------------------------------------------------------------------------------------------------------------
template <int D, int NR>
struct RowsReg {
    static constexpr bool kDynamic = false;
    float v[NR][D];
    HODE_HD void load(int i, float (&o)[D]) const {
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = v[i][d];
    }
    HODE_HD void store(int i, const float (&x)[D]) {
#pragma unroll
        for (int d = 0; d < D; ++d) v[i][d] = x[d];
    }
};
#if HODE_DEVICE_BUILD && defined(__CUDA_ARCH__)
HODE_D float4 ld_row4(const float* p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
HODE_D float2 ld_row2(const float* p) {
    float2 v;
    asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
HODE_D void st_row4(float* p, float a, float b, float c, float d) {
    asm volatile("st.volatile.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(a), "f"(b), "f"(c), "f"(d));
}
HODE_D void st_row2(float* p, float a, float b) {
    asm volatile("st.volatile.shared.v2.f32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(a), "f"(b));
}
#elif HODE_DEVICE_BUILD
HODE_HD float4 ld_row4(const float* p) { return *reinterpret_cast<const float4*>(p); }
HODE_HD float2 ld_row2(const float* p) { return *reinterpret_cast<const float2*>(p); }
HODE_HD void st_row4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
HODE_HD void st_row2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
#endif
template <int D, int NR, int STRIDE = 0>
struct RowsMem {
    static constexpr bool kDynamic = true;
    static constexpr int VEC = (D % 4 == 0) ? 4 : ((D % 2 == 0) ? 2 : 1);
    static constexpr int NC = D / VEC;
    static constexpr int kFloatsPerThread = NR * D;
    float* p;     // this thread's first chunk
    int stride_;  // floats between consecutive chunks of one thread = threads * VEC (used when STRIDE == 0)
    // STRIDE > 0: the stride is a compile-time constant (fixed CTA size): chunk offsets become immediates
    HODE_HD int stride() const { return STRIDE > 0 ? STRIDE : stride_; }
    // The accesses are volatile: with unrolled stage loops the compiler otherwise forwards every stored row to its later
    // loads, i.e. keeps all rows in registers after all (measured: 250 registers or 0.5 KB of spills at D = 12).
    HODE_HD void load(int i, float (&o)[D]) const {
        const float* q = p + (size_t)(i * NC) * stride();
#pragma unroll
        for (int c = 0; c < NC; ++c) {
#if HODE_DEVICE_BUILD
            if (VEC == 4) {
                const float4 t = ld_row4(q + (size_t)c * stride());
                o[4 * c] = t.x; o[4 * c + 1] = t.y; o[4 * c + 2] = t.z; o[4 * c + 3] = t.w;
                continue;
            } else if (VEC == 2) {
                const float2 t = ld_row2(q + (size_t)c * stride());
                o[2 * c] = t.x; o[2 * c + 1] = t.y;
                continue;
            }
#endif
#pragma unroll
            for (int e = 0; e < VEC; ++e) o[c * VEC + e] = q[(size_t)c * stride() + e];
        }
    }
    HODE_HD void store(int i, const float (&x)[D]) {
        float* q = p + (size_t)(i * NC) * stride();
#pragma unroll
        for (int c = 0; c < NC; ++c) {
#if HODE_DEVICE_BUILD
            if (VEC == 4) {
                st_row4(q + (size_t)c * stride(), x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
                continue;
            } else if (VEC == 2) {
                st_row2(q + (size_t)c * stride(), x[2 * c], x[2 * c + 1]);
                continue;
            }
#endif
#pragma unroll
            for (int e = 0; e < VEC; ++e) q[(size_t)c * stride() + e] = x[c * VEC + e];
        }
    }
};

// Stage loops: fully unrolled with compile-time indices (UNROLL: register rows) or rolled.  `fn` receives either a
// std::integral_constant<int, I> or an int; both convert to int.
template <int I>
struct IdxC {
#if HODE_DEVICE_BUILD
    __host__ __device__ constexpr operator int() const { return I; }
#else
    constexpr operator int() const { return I; }
#endif
};
template <bool UNROLL, int LO, int HI, class Fn>  // i = LO .. HI-1 ascending
HODE_HD void stage_up(Fn&& fn) {
    if constexpr (UNROLL) {
        if constexpr (LO < HI) {
            fn(IdxC<LO>{});
            stage_up<true, LO + 1, HI>(fn);
        }
    } else {
#pragma unroll 1
        for (int i = LO; i < HI; ++i) fn(i);
    }
}
template <bool UNROLL, int HI, int LO, class Fn>  // i = HI .. LO descending (inclusive)
HODE_HD void stage_down(Fn&& fn) {
    if constexpr (UNROLL) {
        if constexpr (HI >= LO) {
            fn(IdxC<HI>{});
            stage_down<true, HI - 1, LO>(fn);
        }
    } else {
#pragma unroll 1
        for (int i = HI; i >= LO; --i) fn(i);
    }
}
// Rolled stage loops run the vector field ONCE per trip (one inlined copy), but the stage algebra around it -- which rows
// enter with which tableau coefficients -- is different for every stage: dispatching on the run-time stage index to a
// compile-time one gives each stage its own straight-line code (immediate row offsets, coefficients folded into FMUL
// immediates, no inner loop control).  With run-time inner loops those three were ~35 % of the executed instructions.
template <bool SWITCH, int N, class Fn>
HODE_HD void stage_switch(int i, Fn&& fn) {
    if constexpr (!SWITCH) {
        fn(i);
    } else {
        switch (i) {
            case 0: fn(IdxC<0>{}); break;
            case 1: if constexpr (N > 1) fn(IdxC<1>{}); break;
            case 2: if constexpr (N > 2) fn(IdxC<2>{}); break;
            case 3: if constexpr (N > 3) fn(IdxC<3>{}); break;
            case 4: if constexpr (N > 4) fn(IdxC<4>{}); break;
            case 5: if constexpr (N > 5) fn(IdxC<5>{}); break;
            case 6: if constexpr (N > 6) fn(IdxC<6>{}); break;
            default: break;
        }
    }
}
// j = lo .. hi-1 ascending with run-time bounds inside [0, MAXN): unrolled (bounds fold after inlining) or rolled
template <bool UNROLL, int MAXN, class Fn>
HODE_HD void range_up(int lo, int hi, Fn&& fn) {
    if constexpr (UNROLL) {
#pragma unroll
        for (int j = 0; j < MAXN; ++j)
            if (j >= lo && j < hi) fn(j);
    } else {
#pragma unroll 1
        for (int j = lo; j < hi; ++j) fn(j);
    }
}
template <bool UNROLL, int MAXN, class Fn>  // j = hi-1 .. lo descending
HODE_HD void range_down(int lo, int hi, Fn&& fn) {
    if constexpr (UNROLL) {
#pragma unroll
        for (int j = MAXN - 1; j >= 0; --j)
            if (j >= lo && j < hi) fn(j);
    } else {
#pragma unroll 1
        for (int j = hi - 1; j >= lo; --j) fn(j);
    }
}

}  // namespace hode
