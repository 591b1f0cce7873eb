// hostsim.cpp -- TEST-ONLY thread emulation of the solver kernels (never loaded by the product package).
//
// This container has no GPU, so the per-trajectory bodies of csrc/hode_bodies.cuh (the exact source the sm_100a
// kernels are built from) are also compiled here as plain C++ and driven by host loops / std::thread groups.  It lets
// tests/test_hostsim_*.py check the hand-derived reverse sweeps, tape logic and dense-output emission against the
// oracle without a GPU.  It exports the solver subset of the include/hode.h ABI, taking HOST pointers.
#define HODE_HOSTSIM 1
#define HODE_HAVE_NEURAL 1
#include <algorithm>
#include <barrier>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../hybrid_ode_neurips_2021_b200/csrc/hode_bodies.cuh"
#include "../../hybrid_ode_neurips_2021_b200/csrc/hode_sse.cuh"
#include "../../hybrid_ode_neurips_2021_b200/csrc/hode_real.cuh"

using namespace hode;

namespace {
struct CommNone {
    static constexpr bool kLockstep = false;
    void sum1(float&) {}
    void sum2(float&, float&) {}
    bool any(bool p) { return p; }
    bool all_done(bool done) { return done; }
};
struct CommThreads {
    static constexpr bool kLockstep = false;
    std::barrier<>* bar;
    float* buf;
    int n, tid;
    void sum2(float& a, float& b) {
        buf[2 * tid] = a;
        buf[2 * tid + 1] = b;
        bar->arrive_and_wait();
        float sa = 0.f, sb = 0.f;
        for (int i = 0; i < n; ++i) { sa += buf[2 * i]; sb += buf[2 * i + 1]; }
        bar->arrive_and_wait();
        a = sa; b = sb;
    }
    void sum1(float& a) { float z = 0.f; sum2(a, z); }
    bool any(bool p) { float v = p ? 1.f : 0.f, z = 0.f; sum2(v, z); return v > 0.f; }
    bool all_done(bool done) { return done; }
    // vector reduction of the mixed adjoint norm: out[k] = sum over the group's threads of v[k]
    float* vbuf = nullptr;  // [n][N]
    template <int N>
    void sum_vec(const float (&v)[N], float* out) {
        for (int k = 0; k < N; ++k) vbuf[(size_t)tid * N + k] = v[k];
        bar->arrive_and_wait();
        for (int k = tid; k < N; k += n) {
            float t = 0.f;
            for (int i = 0; i < n; ++i) t += vbuf[(size_t)i * N + k];
            out[k] = t;
        }
        bar->arrive_and_wait();
    }
    void sync() { bar->arrive_and_wait(); }
};
// Emulation of the device's CommPack (several controller groups in one CTA, all threads in lock-step, the reduction's barrier
// doubling as the "did anybody still have work" vote): `n` threads = `groups` x `batch` + padding threads that belong to no group.
struct CommPackThreads {
    static constexpr bool kLockstep = true;
    std::barrier<>* bar;
    float* buf;   // [n] values
    int* idle;    // [n] votes
    int n, tid, first, count;  // my group's threads: first .. first + count - 1 (count = 0: padding thread)
    bool all_idle = false;
    void reduce(float& a, bool idl) {
        buf[tid] = a;
        idle[tid] = idl ? 1 : 0;
        bar->arrive_and_wait();
        float s = 0.f;
        for (int i = 0; i < count; ++i) s += buf[first + i];
        bool all = true;
        for (int i = 0; i < n; ++i) all = all && idle[i] != 0;
        bar->arrive_and_wait();
        all_idle = all;
        a = s;
    }
    void sum1(float& a) { reduce(a, false); }
    void sum1_vote(float& a, bool idl) { reduce(a, idl); }
    void sum2(float& a, float& b) { reduce(a, false); reduce(b, false); }
    bool any(bool p) { float v = p ? 1.f : 0.f; reduce(v, false); return v > 0.f; }
    bool all_done(bool) { return all_idle; }
};
// Row storage of the dopri5 bodies: the device keeps the rows of the D = 12 kernels in shared memory with ROLLED stage
// loops (RowsMem) and in registers with unrolled loops otherwise (RowsReg).  The emulation follows the same policy;
// HODE_HOSTSIM_ROWS=mem / reg forces one of the two code paths for every D (both are tested).
bool rows_in_memory(int D) {
    const char* e = getenv("HODE_HOSTSIM_ROWS");
    if (e && !strcmp(e, "mem")) return true;
    if (e && !strcmp(e, "reg")) return false;
    return D >= 12;
}
template <class F, int NR, class Fn>
void with_rows(Fn&& fn) {
    if (rows_in_memory(F::D)) {
        using RM = RowsMem<F::D, NR>;
        std::vector<float> buf(RM::kFloatsPerThread);
        RM rows{buf.data(), RM::VEC};
        fn(rows, std::true_type{});  // rolled stage loops, like the device's shared-memory variant
    } else {
        RowsReg<F::D, NR> rows;
        fn(rows, std::false_type{});
    }
}

void fill(SolveArgs& a, const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
          const float* dose_t, int64_t stride, const float* params, const int32_t* pset) {
    memset(&a, 0, sizeof(a));
    a.n_groups = n_groups; a.batch = batch; a.dose_amt = dose_amt; a.dose_t = dose_t; a.dose_t_stride = stride;
    a.n_dose = cfg->n_dose; a.params = params; a.pset = pset; a.perturb = cfg->perturb;
    a.rtol_f = (float)cfg->rtol; a.atol_f = (float)cfg->atol; a.safety = cfg->safety; a.ifactor = cfg->ifactor;
    a.dfactor = cfg->dfactor; a.first_step = cfg->first_step; a.max_num_steps = cfg->max_num_steps;
    a.attempt_cap = cfg->attempt_cap; a.per_traj = cfg->controller == HODE_CTRL_TRAJ;
}

template <class F>
std::vector<float> stage(const SolveArgs& a, int64_t g) {
    std::vector<float> sp(F::SP);
    const int set = a.pset ? a.pset[g] : 0;
    F::stage(a.params + (int64_t)set * F::P, sp.data(), 0, 1);
    F::prepare(sp.data());
    return sp;
}
DoseMem dose(const SolveArgs& a, int64_t idx) {
    DoseMem d; d.amt = a.dose_amt[idx]; d.tau = a.dose_t + idx * a.dose_t_stride; d.nd = a.n_dose; return d;
}

template <class F, int M>
void fixed_fwd(const SolveArgs& a) {
    for (int64_t g = 0; g < a.n_groups; ++g) {
        auto sp = stage<F>(a, g);
        for (int64_t b = 0; b < a.batch; ++b) { const int64_t idx = g * a.batch + b; fixed_fwd_traj<F, M>(a, sp.data(), dose(a, idx), idx); }
    }
}
// fused forward + read-out + masked SSE (SseSinkHost): one parameter set, one group
template <class F, int M>
void fixed_fwd_sse(const SolveArgs& a) {
    auto sp = stage<F>(a, 0);
    const int obs = a.sse_obs;
    std::vector<float> gw((size_t)obs * F::D + obs, 0.f);
    double loss = 0.0;
    for (int64_t idx = 0; idx < a.batch; ++idx) {
        SseSinkHost<F::D> sink{&a, a.batch, idx, gw.data(), 0.0};
        fixed_fwd_traj<F, M>(a, sp.data(), dose(a, idx), idx, sink);
        loss += sink.loss;
    }
    *a.sse_loss = (float)(loss * (double)a.sse_inv_norm);
    if (a.sse_grad_w) for (int i = 0; i < obs * F::D; ++i) a.sse_grad_w[i] = gw[i];
    if (a.sse_grad_b) for (int i = 0; i < obs; ++i) a.sse_grad_b[i] = gw[obs * F::D + i];
}
template <class F, int M>
void fixed_bwd(const SolveArgs& a, bool eg) {
    for (int64_t g = 0; g < a.n_groups; ++g) {
        auto sp = stage<F>(a, g);
        std::vector<float> acc(F::P, 0.f);
        for (int64_t b = 0; b < a.batch; ++b) {
            const int64_t idx = g * a.batch + b;
            if (eg) fixed_bwd_traj<F, M, true>(a, sp.data(), dose(a, idx), idx, acc.data());
            else fixed_bwd_traj<F, M, false>(a, sp.data(), dose(a, idx), idx, acc.data());
        }
        const int set = a.pset ? a.pset[g] : 0;
        for (int i = 0; i < F::P; ++i) a.grad_params[(int64_t)set * F::P + i] += acc[i];
    }
}
template <class F, int M>
void fixed_adj(const SolveArgs& a, bool eg) {
    for (int64_t g = 0; g < a.n_groups; ++g) {
        auto sp = stage<F>(a, g);
        std::vector<float> acc(F::P, 0.f);
        for (int64_t b = 0; b < a.batch; ++b) {
            const int64_t idx = g * a.batch + b;
            if (eg) fixed_adj_traj<F, M, true>(a, sp.data(), dose(a, idx), idx, acc.data());
            else fixed_adj_traj<F, M, false>(a, sp.data(), dose(a, idx), idx, acc.data());
        }
        const int set = a.pset ? a.pset[g] : 0;
        for (int i = 0; i < F::P; ++i) a.grad_params[(int64_t)set * F::P + i] += acc[i];
    }
}
template <class F>
void dopri5_fwd(const SolveArgs& a) {
    for (int64_t g = 0; g < a.n_groups; ++g) {
        auto sp = stage<F>(a, g);
        if (a.per_traj) {
            for (int64_t b = 0; b < a.batch; ++b) {
                const int64_t idx = g * a.batch + b;
                CommNone cm;
                with_rows<F, 7>([&](auto& k, auto rolled) { dopri5_fwd_traj<F, decltype(rolled)::value>(a, cm, sp.data(), dose(a, idx), k, idx, true, idx, true, (float)F::D); });
            }
        } else {
            const int n = (int)a.batch;
            std::barrier<> bar(n);
            std::vector<float> buf(2 * n);
            std::vector<std::thread> th;
            for (int b = 0; b < n; ++b)
                th.emplace_back([&, b]() {
                    const int64_t idx = g * a.batch + b;
                    CommThreads cm{&bar, buf.data(), n, b};
                    with_rows<F, 7>([&](auto& k, auto rolled) { dopri5_fwd_traj<F, decltype(rolled)::value>(a, cm, sp.data(), dose(a, idx), k, idx, true, g, b == 0, (float)(a.batch * F::D)); });
                });
            for (auto& t : th) t.join();
        }
    }
}
// the device's packed launch shape (dopri5_fwd_pack_kernel): `gpc` groups per emulated CTA plus two padding threads, lock-step
template <class F>
void dopri5_fwd_packed(const SolveArgs& a, int gpc) {
    auto sp = stage<F>(a, 0);
    const int bsz = (int)a.batch;
    for (int64_t g0 = 0; g0 < a.n_groups; g0 += gpc) {
        const int n = gpc * bsz + 2;  // a last CTA with fewer groups has more padding threads, like on the device
        std::barrier<> bar(n);
        std::vector<float> buf(n);
        std::vector<int> idle(n);
        std::vector<std::thread> th;
        for (int t = 0; t < n; ++t)
            th.emplace_back([&, t]() {
                const int gl = t / bsz;
                const int64_t g = g0 + gl;
                const bool valid = gl < gpc && g < a.n_groups;
                const int rel = t - gl * bsz;
                const int64_t idx = valid ? g * a.batch + rel : 0;
                CommPackThreads cm{&bar, buf.data(), idle.data(), n, t, gl * bsz, valid ? bsz : 0};
                with_rows<F, 7>([&](auto& k, auto rolled) { dopri5_fwd_traj<F, decltype(rolled)::value>(a, cm, sp.data(), dose(a, idx), k, idx, valid, valid ? g : 0, valid && rel == 0, (float)(a.batch * F::D)); });
            });
        for (auto& t : th) t.join();
    }
}
template <class F>
void dopri5_bwd(const SolveArgs& a, bool eg) {
    for (int64_t g = 0; g < a.n_groups; ++g) {
        auto sp = stage<F>(a, g);
        std::vector<float> acc(F::P, 0.f);
        for (int64_t b = 0; b < a.batch; ++b) {
            const int64_t idx = g * a.batch + b;
            const int64_t ctrl = a.per_traj ? idx : g;
            // nloop = the longest tape of the launch: exercises the idle (end-aligned) iterations of the device's
            // warp-cooperative path as well
            int nloop = 0;
            const int64_t n_ctrl = a.per_traj ? a.n_groups * a.batch : a.n_groups;
            for (int64_t c = 0; c < n_ctrl; ++c) nloop = std::max(nloop, (int)a.stats[c].accepted);
            with_rows<F, 9>([&](auto& R, auto rolled) {
                constexpr bool RL = decltype(rolled)::value;
                if (eg) dopri5_bwd_traj<F, true, RL>(a, sp.data(), dose(a, idx), R, idx, ctrl, acc.data(), true, nloop);
                else dopri5_bwd_traj<F, false, RL>(a, sp.data(), dose(a, idx), R, idx, ctrl, acc.data(), true, nloop);
            });
        }
        const int set = a.pset ? a.pset[g] : 0;
        for (int i = 0; i < F::P; ++i) a.grad_params[(int64_t)set * F::P + i] += acc[i];
    }
}

// adaptive continuous adjoint: 14 rows in memory, rolled stage loops (the device's only variant)
template <class F>
void dopri5_adj(const SolveArgs& a, bool eg) {
    using RM = RowsMem<F::D, 14>;
    for (int64_t g = 0; g < a.n_groups; ++g) {
        auto sp = stage<F>(a, g);
        std::vector<float> acc(F::P, 0.f);
        auto one = [&](auto& cm, int64_t idx, int64_t ctrl, bool leader, float count, float* accp) {
            std::vector<float> buf(RM::kFloatsPerThread);
            RM R{buf.data(), RM::VEC};
            if (eg) dopri5_adj_traj<F, true, true>(a, cm, sp.data(), dose(a, idx), R, idx, true, ctrl, leader, count, accp);
            else dopri5_adj_traj<F, false, true>(a, cm, sp.data(), dose(a, idx), R, idx, true, ctrl, leader, count, accp);
        };
        if constexpr (F::kAccInRegs && F::D <= 8) {
            if (a.adj_mixed && !a.per_traj) {  // torchdiffeq's default mixed norm (dopri5_adj_mixed_traj)
                const int n = (int)a.batch;
                std::barrier<> bar(n);
                std::vector<float> buf(2 * n), vbuf((size_t)n * 2 * F::P), pg(F::P, 0.f), pq(2 * F::P, 0.f), pr(4, 0.f);
                std::vector<std::thread> th;
                for (int b = 0; b < n; ++b)
                    th.emplace_back([&, b]() {
                        CommThreads cm{&bar, buf.data(), n, b};
                        cm.vbuf = vbuf.data();
                        ParamCtl pc{pg.data(), pq.data(), pr.data()};
                        std::vector<float> rb(RM::kFloatsPerThread);
                        RM R{rb.data(), RM::VEC};
                        const int64_t idx = g * a.batch + b;
                        if (eg) dopri5_adj_mixed_traj<F, true>(a, cm, sp.data(), dose(a, idx), R, idx, true, g, b == 0, (float)(a.batch * F::D), pc, b, n);
                        else dopri5_adj_mixed_traj<F, false>(a, cm, sp.data(), dose(a, idx), R, idx, true, g, b == 0, (float)(a.batch * F::D), pc, b, n);
                    });
                for (auto& t : th) t.join();
                const int set = a.pset ? a.pset[g] : 0;
                for (int i = 0; i < F::P; ++i) a.grad_params[(int64_t)set * F::P + i] += pg[i];
                continue;
            }
        }
        if (a.per_traj) {
            for (int64_t b = 0; b < a.batch; ++b) {
                const int64_t idx = g * a.batch + b;
                CommNone cm;
                one(cm, idx, idx, true, (float)F::D, acc.data());
            }
        } else {
            const int n = (int)a.batch;
            std::barrier<> bar(n);
            std::vector<float> buf(2 * n);
            std::vector<std::vector<float>> accs(n, std::vector<float>(F::P, 0.f));
            std::vector<std::thread> th;
            for (int b = 0; b < n; ++b)
                th.emplace_back([&, b]() {
                    CommThreads cm{&bar, buf.data(), n, b};
                    one(cm, g * a.batch + b, g, b == 0, (float)(a.batch * F::D), accs[b].data());
                });
            for (auto& t : th) t.join();
            for (int b = 0; b < n; ++b)
                for (int i = 0; i < F::P; ++i) acc[i] += accs[b][i];
        }
        const int set = a.pset ? a.pset[g] : 0;
        for (int i = 0; i < F::P; ++i) a.grad_params[(int64_t)set * F::P + i] += acc[i];
    }
}

enum Op { FF, FB, DF, DB, FA, FS, DA };
template <class F>
int run(Op op, const hode_cfg& cfg, const SolveArgs& a) {
    const bool eg = cfg.expert_grads != 0;
    switch (op) {
        case FF:
            if (cfg.method == HODE_EULER) fixed_fwd<F, M_EULER>(a);
            else if (cfg.method == HODE_MIDPOINT) fixed_fwd<F, M_MIDPOINT>(a);
            else fixed_fwd<F, M_RK4_38>(a);
            return 0;
        case FB:
            if (cfg.method == HODE_EULER) fixed_bwd<F, M_EULER>(a, eg);
            else if (cfg.method == HODE_MIDPOINT) fixed_bwd<F, M_MIDPOINT>(a, eg);
            else fixed_bwd<F, M_RK4_38>(a, eg);
            return 0;
        case FA:
            if (cfg.method == HODE_EULER) fixed_adj<F, M_EULER>(a, eg);
            else if (cfg.method == HODE_MIDPOINT) fixed_adj<F, M_MIDPOINT>(a, eg);
            else fixed_adj<F, M_RK4_38>(a, eg);
            return 0;
        case FS:
            if (cfg.method == HODE_EULER) fixed_fwd_sse<F, M_EULER>(a);
            else if (cfg.method == HODE_MIDPOINT) fixed_fwd_sse<F, M_MIDPOINT>(a);
            else fixed_fwd_sse<F, M_RK4_38>(a);
            return 0;
        case DF: {
            const char* pk = getenv("HODE_HOSTSIM_PACK");  // test switch: emulate the packed launch shape with this many groups per CTA
            if (pk != nullptr && !a.per_traj && a.pset == nullptr && atoi(pk) > 0) dopri5_fwd_packed<F>(a, atoi(pk));
            else dopri5_fwd<F>(a);
            return 0;
        }
        case DB: dopri5_bwd<F>(a, eg); return 0;
        case DA: dopri5_adj<F>(a, eg); return 0;
    }
    return -1;
}
int dispatch(Op op, const hode_cfg& cfg, const SolveArgs& a) {
    if (cfg.field == HODE_FIELD_ROCHE) {
        const bool h2 = (cfg.flags & HODE_FLAG_HILL2) != 0;
        if (cfg.flags & HODE_FLAG_ABLATE) {
            switch (cfg.latent_dim) {
                case 4: return run<Roche<4, true, true>>(op, cfg, a);
                case 6: return run<Roche<6, true, true>>(op, cfg, a);
                case 8: return run<Roche<8, true, true>>(op, cfg, a);
                case 12: return run<Roche<12, true, true>>(op, cfg, a);
            }
        }
        switch (cfg.latent_dim) {
            case 4: return h2 ? run<Roche<4, true>>(op, cfg, a) : run<Roche<4>>(op, cfg, a);
            case 6: return h2 ? run<Roche<6, true>>(op, cfg, a) : run<Roche<6>>(op, cfg, a);
            case 8: return h2 ? run<Roche<8, true>>(op, cfg, a) : run<Roche<8>>(op, cfg, a);
            case 12: return h2 ? run<Roche<12, true>>(op, cfg, a) : run<Roche<12>>(op, cfg, a);
        }
    }
#ifdef HODE_HAVE_NEURAL
    if (cfg.field == HODE_FIELD_NEURAL) {
        switch (cfg.latent_dim) {
            case 4: return run<Neural<4>>(op, cfg, a);
            case 6: return run<Neural<6>>(op, cfg, a);
            case 8: return run<Neural<8>>(op, cfg, a);
            case 12: return run<Neural<12>>(op, cfg, a);
        }
    }
#endif
    return HODE_ERR_UNSUPPORTED;
}
int64_t pcount(const hode_cfg* cfg) {
    const int64_t d = cfg->latent_dim;
    if (cfg->field == HODE_FIELD_ROCHE) return 13 + (d - 4) * d + (d - 4) + ((cfg->flags & HODE_FLAG_ABLATE) ? 2 : 0);
    return 1 + 10 * d * (d + 1) + 10 * d + d * 10 * d + d;
}
}  // namespace

extern "C" {
int32_t hode_abi_version(void) { return HODE_ABI_VERSION; }
const char* hode_last_error(void) { return "hostsim"; }
int64_t hode_param_count(const hode_cfg* cfg) { return pcount(cfg); }
int64_t hode_dopri5_max_batch(const hode_cfg*) { return 256; }

int32_t hode_fixed_fwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* y0, const float* dose_amt,
                       const float* dose_t, int64_t dose_t_stride, const float* params, const int32_t* pset,
                       const float* grid, int32_t n_grid, const float* t_eval, int32_t n_t, float* h_out, float* tape,
                       void*) {
    SolveArgs a; fill(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, pset);
    a.y0 = y0; a.grid = grid; a.n_grid = n_grid; a.t_eval_f = t_eval; a.n_t = n_t; a.h_out = h_out; a.tape_y = tape;
    return dispatch(FF, *cfg, a);
}
int32_t hode_fixed_fwd_sse_supported(const hode_cfg* cfg, int32_t obs, int32_t n_param_sets) {
    // same rule as the device library (csrc/hode_api.cu)
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
                           float* grad_w, float* grad_b, void*) {
    if (!hode_fixed_fwd_sse_supported(cfg, obs, 1)) return HODE_ERR_UNSUPPORTED;
    SolveArgs a; fill(a, cfg, 1, n_traj, dose_amt, dose_t, dose_t_stride, params, nullptr);
    a.y0 = y0; a.grid = grid; a.n_grid = n_grid; a.t_eval_f = t_eval; a.n_t = n_t; a.h_out = h_out; a.tape_y = tape;
    a.sse_x = x; a.sse_mask = mask; a.sse_w = W; a.sse_b = b; a.sse_obs = obs;
    a.sse_scale = (float)(-2.0 / n_norm); a.sse_inv_norm = (float)(1.0 / n_norm);
    a.sse_loss = loss; a.sse_grad_h = grad_h; a.sse_grad_w = grad_w; a.sse_grad_b = grad_b;
    *loss = 0.f;
    return dispatch(FS, *cfg, a);
}
int32_t hode_fixed_bwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                       const float* dose_t, int64_t dose_t_stride, const float* params, const int32_t* pset,
                       int32_t n_param_sets, const float* grid, int32_t n_grid, const float* t_eval, int32_t n_t,
                       const float* grad_h, const float* tape, float* grad_y0, float* grad_params, void*) {
    SolveArgs a; fill(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, pset);
    a.n_param_sets = n_param_sets; a.grid = grid; a.n_grid = n_grid; a.t_eval_f = t_eval; a.n_t = n_t;
    a.grad_h = grad_h; a.tape_y = const_cast<float*>(tape); a.grad_y0 = grad_y0; a.grad_params = grad_params;
    memset(grad_params, 0, sizeof(float) * pcount(cfg) * n_param_sets);
    return dispatch(FB, *cfg, a);
}
int32_t hode_fixed_adjoint(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                           const float* dose_t, int64_t dose_t_stride, const float* params, const int32_t* pset,
                           int32_t n_param_sets, const float* adj_grid, int32_t n_adj_grid, const int32_t* adj_count,
                           int32_t n_t, const float* h, const float* grad_h, float* grad_y0, float* grad_params, void*) {
    SolveArgs a; fill(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, pset);
    a.n_param_sets = n_param_sets; a.grid = adj_grid; a.n_grid = n_adj_grid; a.adj_cnt = adj_count; a.n_t = n_t;
    a.h_out = const_cast<float*>(h); a.grad_h = grad_h; a.grad_y0 = grad_y0; a.grad_params = grad_params;
    memset(grad_params, 0, sizeof(float) * pcount(cfg) * n_param_sets);
    return dispatch(FA, *cfg, a);
}
int32_t hode_dopri5_fwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* y0, const float* dose_amt,
                        const float* dose_t, int64_t dose_t_stride, const float* params, const int32_t* pset,
                        const double* t_eval, int32_t n_t, float* h_out, double* tape_t, float* tape_y,
                        int32_t tape_capacity, hode_stats* stats, void*) {
    SolveArgs a; fill(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, pset);
    a.y0 = y0; a.t_eval_d = t_eval; a.n_t = n_t; a.h_out = h_out; a.tape_t = tape_t; a.tape_y = tape_y;
    a.tape_cap = tape_capacity; a.stats = stats;
    return dispatch(DF, *cfg, a);
}
int32_t hode_dopri5_bwd(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                        const float* dose_t, int64_t dose_t_stride, const float* params, const int32_t* pset,
                        int32_t n_param_sets, const double* t_eval, int32_t n_t, const float* grad_h,
                        const double* tape_t, const float* tape_y, int32_t tape_capacity, const hode_stats* stats,
                        float* grad_y0, float* grad_params, void*) {
    SolveArgs a; fill(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, pset);
    a.n_param_sets = n_param_sets; a.t_eval_d = t_eval; a.n_t = n_t; a.grad_h = grad_h;
    a.tape_t = const_cast<double*>(tape_t); a.tape_y = const_cast<float*>(tape_y); a.tape_cap = tape_capacity;
    a.stats = const_cast<hode_stats*>(stats); a.grad_y0 = grad_y0; a.grad_params = grad_params;
    memset(grad_params, 0, sizeof(float) * pcount(cfg) * n_param_sets);
    return dispatch(DB, *cfg, a);
}
int32_t hode_dopri5_adjoint(const hode_cfg* cfg, int64_t n_groups, int64_t batch, const float* dose_amt,
                            const float* dose_t, int64_t dose_t_stride, const float* params, const int32_t* pset,
                            int32_t n_param_sets, const double* t_eval, int32_t n_t, const float* h, const float* grad_h,
                            float* grad_y0, float* grad_params, hode_stats* stats, void*) {
    const bool mixed = !(cfg->flags & HODE_FLAG_ADJ_SEMINORM);
    if (mixed && (cfg->controller != HODE_CTRL_BATCH || cfg->field != HODE_FIELD_ROCHE || cfg->latent_dim > 8)) return HODE_ERR_UNSUPPORTED;
    SolveArgs a; fill(a, cfg, n_groups, batch, dose_amt, dose_t, dose_t_stride, params, pset);
    a.n_param_sets = n_param_sets; a.t_eval_d = t_eval; a.n_t = n_t; a.h_out = const_cast<float*>(h); a.grad_h = grad_h;
    a.stats = stats; a.grad_y0 = grad_y0; a.grad_params = grad_params;
    a.adj_mixed = mixed ? 1 : 0;
    memset(grad_params, 0, sizeof(float) * pcount(cfg) * n_param_sets);
    return dispatch(DA, *cfg, a);
}
// ---- real-data fields (csrc/hode_real.cuh) --------------------------------------------------------------------------
}  // extern "C"
namespace {
template <class F, bool TWO>
int real_run(bool bwd, int method, int hidden, int P, const SolveArgs& a, const float* tab, int T) {
    const int64_t n = a.n_groups * a.batch;
    typename F::Params sp{a.params, hidden};
    std::vector<float> acc(P, 0.f);
    for (int64_t idx = 0; idx < n; ++idx) {
        DoseTab d;
        d.s = tab + idx; d.s1 = TWO ? tab + (int64_t)(T + 1) * n + idx : nullptr; d.stride = n; d.T = T;
        if (!bwd) {
            if (method == HODE_EULER) fixed_fwd_traj<F, M_EULER>(a, sp, d, idx);
            else if (method == HODE_MIDPOINT) fixed_fwd_traj<F, M_MIDPOINT>(a, sp, d, idx);
            else fixed_fwd_traj<F, M_RK4_38>(a, sp, d, idx);
        } else {
            if (method == HODE_EULER) fixed_bwd_traj<F, M_EULER, true>(a, sp, d, idx, acc.data());
            else if (method == HODE_MIDPOINT) fixed_bwd_traj<F, M_MIDPOINT, true>(a, sp, d, idx, acc.data());
            else fixed_bwd_traj<F, M_RK4_38, true>(a, sp, d, idx, acc.data());
        }
    }
    if (bwd) for (int i = 0; i < P; ++i) a.grad_params[i] = acc[i];
    return 0;
}
int real_pcount(int field, int Z, int H) {
    if (H < 1 || H > 64) return -1;
    if (field == HODE_FIELD_ROCHE_REAL) return Z == 4 ? RocheReal<4>::p_count(H) : Z == 20 ? RocheReal<20>::p_count(H) : -1;
    if (field == HODE_FIELD_NEURAL_REAL) return Z == 4 ? NeuralReal<4, false>::p_count(H) : Z == 20 ? NeuralReal<20, false>::p_count(H) : -1;
    if (field == HODE_FIELD_NEURAL_REAL_2ND) return Z == 8 ? NeuralReal<8, true>::p_count(H) : Z == 40 ? NeuralReal<40, true>::p_count(H) : -1;
    return -1;
}
int real_dispatch(bool bwd, int field, int Z, int hidden, int method, const SolveArgs& a, const float* tab, int T) {
    const int P = real_pcount(field, Z, hidden);
    if (P < 0) return HODE_ERR_UNSUPPORTED;
    if (field == HODE_FIELD_ROCHE_REAL) {
        if (Z == 4) return real_run<RocheReal<4>, true>(bwd, method, hidden, P, a, tab, T);
        return real_run<RocheReal<20>, true>(bwd, method, hidden, P, a, tab, T);
    }
    if (field == HODE_FIELD_NEURAL_REAL) {
        if (Z == 4) return real_run<NeuralReal<4, false>, false>(bwd, method, hidden, P, a, tab, T);
        return real_run<NeuralReal<20, false>, false>(bwd, method, hidden, P, a, tab, T);
    }
    if (Z == 8) return real_run<NeuralReal<8, true>, false>(bwd, method, hidden, P, a, tab, T);
    return real_run<NeuralReal<40, true>, false>(bwd, method, hidden, P, a, tab, T);
}
}  // namespace
extern "C" {
int64_t hode_real_param_count(int32_t field, int32_t latent_dim, int32_t hidden) { return real_pcount(field, latent_dim, hidden); }
int32_t hode_real_dose_tables(int32_t field, const float* action, int64_t stride_t, int64_t stride_b, int32_t T,
                              int64_t n_traj, const float* params, float* tab, void*) {
    const int kind = field == HODE_FIELD_ROCHE_REAL ? 0 : 1;
    for (int64_t b = 0; b < n_traj; ++b)
        real_dose_table_column(kind, action + b * stride_b, stride_t, T, n_traj, kind == 0 ? params[1] : 0.f, tab, b);
    return 0;
}
int32_t hode_real_fixed_fwd(int32_t field, int32_t latent_dim, int32_t hidden, int32_t method, int32_t perturb,
                            int64_t n_traj, const float* y0, const float* tab, int32_t T, const float* params,
                            const float* grid, int32_t n_grid, const float* t_eval, int32_t n_t, float* h_out,
                            float* tape, void*) {
    SolveArgs a; memset(&a, 0, sizeof(a));
    a.n_groups = 1; a.batch = n_traj; a.params = params; a.perturb = perturb; a.grid = grid; a.n_grid = n_grid;
    a.t_eval_f = t_eval; a.n_t = n_t; a.y0 = y0; a.h_out = h_out; a.tape_y = tape;
    return real_dispatch(false, field, latent_dim, hidden, method, a, tab, T);
}
int32_t hode_real_fixed_bwd(int32_t field, int32_t latent_dim, int32_t hidden, int32_t method, int32_t perturb,
                            int64_t n_traj, const float* tab, int32_t T, const float* params, const float* grid,
                            int32_t n_grid, const float* t_eval, int32_t n_t, const float* grad_h, const float* tape,
                            float* grad_y0, float* grad_params, void*) {
    SolveArgs a; memset(&a, 0, sizeof(a));
    a.n_groups = 1; a.batch = n_traj; a.params = params; a.perturb = perturb; a.grid = grid; a.n_grid = n_grid;
    a.t_eval_f = t_eval; a.n_t = n_t; a.grad_h = grad_h; a.tape_y = const_cast<float*>(tape); a.grad_y0 = grad_y0;
    a.grad_params = grad_params;
    return real_dispatch(true, field, latent_dim, hidden, method, a, tab, T);
}
}
