// hode_launch.cuh -- sm_100a kernels wrapping the per-trajectory bodies, and their launchers.
//
// Mapping: one thread = one trajectory, whole solve in one launch, state/stages in registers; the parameter set of the
// CTA's group is staged once into shared memory (broadcast LDS on every use); observations of the solution are
// written with 16/8-byte vector stores, D*4 contiguous bytes per lane (full 32 B sectors).  A CTA never spans two
// groups, so a CTA has exactly one parameter set and (dopri5, batch-coupled) one controller.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include <mutex>

#include "hode_bodies.cuh"
#include "hode_sse.cuh"

#ifndef HODE_FWD_MINBLOCKS
#define HODE_FWD_MINBLOCKS 1
#endif
#ifndef HODE_BWD_MINBLOCKS
#define HODE_BWD_MINBLOCKS 1
#endif
#ifndef HODE_D5_FWD_MINBLOCKS
#define HODE_D5_FWD_MINBLOCKS 4
#endif
#ifndef HODE_D5_FWD_MINBLOCKS_REG
#define HODE_D5_FWD_MINBLOCKS_REG 4
#endif
#ifndef HODE_SSE_MINBLOCKS
#define HODE_SSE_MINBLOCKS 4
#endif
#ifndef HODE_DOPRI5_MAX_THREADS
#define HODE_DOPRI5_MAX_THREADS 512
#endif

namespace hode {

__host__ __device__ constexpr int round4(int n) { return (n + 3) / 4 * 4; }

// ---- group operations of the dopri5 controller: a sum over the controller group (batch-coupled: the CTA or a lane
//      segment; per-trajectory: nothing) that is bit-identical in every thread of the group, and any() -------------------
struct CommNone {
    static constexpr bool kLockstep = false;
    __device__ __forceinline__ void sum1(float&) {}
    __device__ __forceinline__ void sum2(float&, float&) {}
    __device__ __forceinline__ bool any(bool p) { return p; }
    __device__ __forceinline__ bool all_done(bool done) { return done; }
};
struct CommCta {
    static constexpr bool kLockstep = false;
    float* red;  // 2 buffers x (2 * 32) floats of shared memory, used alternately: ONE barrier per reduction
    int nwarps;
    int phase = 0;
    __device__ __forceinline__ void sum2(float& a, float& b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (nwarps > 1) {
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            float* r = red + phase * 64;
            if (lane == 0) { r[2 * wid] = a; r[2 * wid + 1] = b; }
            __syncthreads();
            float sa = 0.0f, sb = 0.0f;
            for (int i = 0; i < nwarps; ++i) { sa += r[2 * i]; sb += r[2 * i + 1]; }
            // no second barrier: the next reduction writes the OTHER buffer, and the one after that is separated from these
            // reads by the next reduction's barrier
            phase ^= 1;
            a = sa; b = sb;
        }
    }
    __device__ __forceinline__ void sum1(float& a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (nwarps > 1) {
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            float* r = red + phase * 64;
            if (lane == 0) r[wid] = a;
            __syncthreads();
            float sa = 0.0f;
            for (int i = 0; i < nwarps; ++i) sa += r[i];
            phase ^= 1;
            a = sa;
        }
    }
    __device__ __forceinline__ bool any(bool p) {
        if (nwarps > 1) return __syncthreads_or(p ? 1 : 0) != 0;
        return __any_sync(0xffffffffu, p) != 0;
    }
    __device__ __forceinline__ bool all_done(bool done) { return done; }  // CTA-uniform by construction
    // vector reduction for the mixed adjoint norm: out[k] = sum over the CTA of v[k], visible to every thread on return
    float* vbuf = nullptr;  // [nwarps][N]
    template <int N>
    __device__ __forceinline__ void sum_vec(const float (&v)[N], float* out) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        auto one = [&](int k) {
            float x = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) vbuf[wid * N + k] = x;
        };
        if constexpr (N <= 64) {  // compile-time indices: `v` can stay in registers
#pragma unroll
            for (int k = 0; k < N; ++k) one(k);
        } else {
#pragma unroll 1
            for (int k = 0; k < N; ++k) one(k);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            float t = 0.0f;
            for (int w = 0; w < nwarps; ++w) t += vbuf[w * N + k];
            out[k] = t;
        }
        __syncthreads();
    }
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

// ---------------------------------------------------------------------------------------------------------------
// Constant-bank parameters.  When a launch has ONE parameter set (every reference call site: model.py:1116), the staged
// RocheODE parameters are copied (device to device, stream-ordered) into this __constant__ array and the kernels read
// them as c[bank][imm] operands of FFMA/FMUL: no shared-memory loads and, above all, no registers (36-117 floats that
// ptxas otherwise keeps live across the whole solve).  Launches with several parameter sets stage into shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kConstParamMax = 384;  // >= Roche<12, *, true>::SP
static __constant__ __align__(16) float c_params[kConstParamMax];
static __device__ __align__(16) float g_param_stage[kConstParamMax];
static_assert(Roche<12, true, true>::SP <= kConstParamMax, "constant-bank parameter array too small");

struct ParamConst {
    __device__ __forceinline__ float operator[](int i) const { return c_params[i]; }
};

template <class F>
__global__ void __launch_bounds__(128) prep_params_kernel(const float* __restrict__ src) {
    F::stage(src, g_param_stage, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) F::prepare(g_param_stage);
}

// The constant array is per-device state shared by every stream: a lease serialises its users (host mutex for the
// bookkeeping, an event so that a launch on another stream waits until the previous user's kernel has finished).
struct ConstParamLease {
    static constexpr int kMaxDev = 64;
    struct PerDevice { cudaEvent_t ev = nullptr; cudaStream_t last = nullptr; bool used = false; };
    static std::mutex& mu() { static std::mutex m; return m; }
    static PerDevice* slots() { static PerDevice s[kMaxDev]; return s; }
    std::unique_lock<std::mutex> lock;
    PerDevice* pd = nullptr;
    cudaStream_t st;
    int rc = 0;
    template <class F>
    ConstParamLease(F*, const float* params, cudaStream_t stream) : lock(mu()), st(stream) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess || dev < 0 || dev >= kMaxDev) { rc = (int)(e != cudaSuccess ? e : cudaErrorInvalidDevice); return; }
        pd = &slots()[dev];
        if (pd->ev == nullptr) {
            e = cudaEventCreateWithFlags(&pd->ev, cudaEventDisableTiming);
            if (e != cudaSuccess) { rc = (int)e; return; }
        }
        if (pd->used && pd->last != st) {
            e = cudaStreamWaitEvent(st, pd->ev, 0);
            if (e != cudaSuccess) { rc = (int)e; return; }
        }
        prep_params_kernel<F><<<1, 128, 0, st>>>(params);
        e = cudaGetLastError();
        if (e != cudaSuccess) { rc = (int)e; return; }
        void* stage_ptr = nullptr;
        e = cudaGetSymbolAddress(&stage_ptr, g_param_stage);
        if (e != cudaSuccess) { rc = (int)e; return; }
        e = cudaMemcpyToSymbolAsync(c_params, stage_ptr, sizeof(float) * F::SP, 0, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { rc = (int)e; return; }
    }
    // call after the consumer kernel has been enqueued
    void done() {
        if (pd != nullptr && rc == 0) {
            cudaEventRecord(pd->ev, st);
            pd->last = st;
            pd->used = true;
        }
    }
};

inline bool const_params_enabled() {
    static const bool on = getenv("HODE_NO_CONST_PARAMS") == nullptr;
    return on;
}
inline bool pack_enabled() {  // HODE_NO_PACK=1: one group per CTA for mid-size groups (A/B measurements, tests of both paths)
    static const bool on = getenv("HODE_NO_PACK") == nullptr;
    return on;
}
template <class F>
inline bool use_const_params(const SolveArgs& a) { return F::kConstBank && a.pset == nullptr && const_params_enabled(); }

// several small groups share one warp: a group is a segment of `size` consecutive lanes with its own accept / reject
// sequence (independent thread scheduling lets the segments of a warp diverge).  The group sum goes through a warp-private
// shared-memory buffer -- one store, one __syncwarp over the segment's lanes, round_up4(size) / 4 vector loads: every lane
// adds the segment's values in the same order, so the sum (and the accept decision) is bit-identical across the group.
// (Per-lane __shfl_sync with a run-time mask cost a MATCH + REDUX + VOTE + WARPSYNC sequence per shuffle: ~120
// instructions per reduction of a 10-lane group.)  Two buffers are used alternately, so one __syncwarp per reduction.
struct CommSeg {
    static constexpr bool kLockstep = false;
    unsigned mask;  // lanes of this segment
    int rel;        // lane index inside the segment
    int nvec;       // round_up4(size) / 4
    float* buf;     // this segment's slots: [2 phases][16] floats, padding slots stay zero
    int phase = 0;
    static constexpr int kFloatsPerWarp = 2 * 8 * 16;  // up to 8 segments (size >= 4 ... the launcher uses size <= 16)
    __device__ __forceinline__ void sum1(float& a) {
        float* b = buf + phase * 16;
        b[rel] = a;
        __syncwarp(mask);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        float4 q = b4[0];
        float s = ((q.x + q.y) + q.z) + q.w;
        if (nvec > 1) { q = b4[1]; s = (((s + q.x) + q.y) + q.z) + q.w; }
        if (nvec > 2) { q = b4[2]; s = (((s + q.x) + q.y) + q.z) + q.w; }
        if (nvec > 3) { q = b4[3]; s = (((s + q.x) + q.y) + q.z) + q.w; }
        phase ^= 1;
        a = s;
    }
    __device__ __forceinline__ void sum2(float& a, float& b) { sum1(a); sum1(b); }
    __device__ __forceinline__ bool any(bool p) { return (__ballot_sync(mask, p) & mask) != 0u; }
    // re-converges every controller of the warp (`part` = the warp's participating lanes) and tells whether all are done
    unsigned part;
    __device__ __forceinline__ bool all_done(bool done) { return __all_sync(part, done) != 0; }
};

// A group LARGER than one CTA (batch > 512; torchdiffeq's norm is over the whole [B, D] tensor at any B, model.py:1116): the
// group's CTAs form a thread-block CLUSTER (<= 8 CTAs = 4 096 trajectories).  Per reduction: CTA-level sum as above, then the
// first `nrank` threads of every CTA store that sum into slot [own rank] of EVERY CTA's shared memory (distributed shared
// memory), one cluster barrier, and every thread adds the slots in rank order -- bit-identical across the whole group, so the
// accept / reject decisions (and the loop trip count, hence the barrier count) are cluster-uniform by construction.
struct CommCluster {
    static constexpr bool kLockstep = false;
    static constexpr int kMaxCtas = 8, kFloats = 2 * kMaxCtas;
    CommCta cta;
    float* slots;  // [2 phases][kMaxCtas] in this CTA's shared memory
    unsigned rank, nrank;
    int phase = 0;
    __device__ __forceinline__ void sum1(float& a) {
        cta.sum1(a);
        namespace cg = cooperative_groups;
        cg::cluster_group cl = cg::this_cluster();
        float* mine = slots + phase * kMaxCtas;
        if (threadIdx.x < nrank) cl.map_shared_rank(mine, threadIdx.x)[rank] = a;
        cl.sync();  // release / acquire: the remote stores are visible after it
        float s = 0.0f;
        for (unsigned i = 0; i < nrank; ++i) s += mine[i];
        phase ^= 1;  // as in CommCta: the buffer written two reductions from now is separated from these reads by the next barrier
        a = s;
    }
    __device__ __forceinline__ void sum2(float& a, float& b) { sum1(a); sum1(b); }
    __device__ __forceinline__ bool any(bool p) { float f = p ? 1.0f : 0.0f; sum1(f); return f > 0.0f; }
    __device__ __forceinline__ bool all_done(bool done) { return done; }  // cluster-uniform by construction
};

// Mid-size groups PACKED into one CTA: floor(256 / batch) groups of `batch` consecutive threads each, group boundaries
// anywhere inside a warp.  (One group per CTA leaves 14 of 64 lanes idle at the reference's batch of 50,
// run_simulation.py:30; 5 groups in 256 threads use 250.)  Group sum in two levels: per warp, one masked butterfly per group
// that overlaps the warp (at most 3 for batch >= 17) -> lane 0 stores the partial in the group's slot of that warp -> ONE CTA
// barrier -> every thread adds its group's partials in warp order (bit-identical across the group).  All threads of the CTA
// walk the attempts in LOCK-STEP -- a finished or failed group keeps executing (results discarded) so that every thread meets
// every barrier -- and the barrier doubles as the vote that ends the loop: sum1_vote(.., idle) returns in `all_idle` whether
// no thread of the CTA had work in this attempt.
struct CommPack {
    static constexpr bool kLockstep = true;
    static constexpr int kThreads = 256, kW = 8, kG = 16;  // largest CTA; warps per CTA (>= warps a group can span); groups per CTA (batch >= 17 -> <= 15)
    // CTA size for groups of `batch`: the smallest multiple of 32 (<= 256) that fills >= 93 % of its lanes with whole groups --
    // small CTAs keep the lock-step barrier cheap and the wait for a CTA's slowest group short -- else the best-filled one.
    static int threads_for(int batch) {
        static const int forced = getenv("HODE_PACK_THREADS") ? atoi(getenv("HODE_PACK_THREADS")) : 0;  // A/B measurements
        if (forced >= 32 && forced <= kThreads && forced % 32 == 0 && forced >= batch) return forced;
        int best = 0;
        double best_u = 0.0;
        for (int t = 32; t <= kThreads; t += 32) {
            const double u = (double)(t / batch * batch) / t;
            if (u >= 0.93) return t;
            if (u > best_u) { best_u = u; best = t; }
        }
        return best;
    }
    static constexpr int kFloats = 2 * kG * kW;
    float* red;        // [2 phases][kG][kW]
    int batch;
    int g;             // this thread's group inside the CTA; kG - 1 .. for padding threads: a group nobody reduces into
    int g_lo, g_hi;    // groups overlapping this warp (warp-uniform; g_lo > g_hi: none)
    int nw;            // warps my group spans (0 for padding threads)
    int phase = 0;
    bool all_idle = false;
    __device__ __forceinline__ void reduce(float& a, bool idle) {
        float* r = red + phase * (kG * kW);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int gk = g_lo + k;
            if (gk <= g_hi) {  // warp-uniform
                float v = (g == gk) ? a : 0.0f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) r[gk * kW + (wid - ((gk * batch) >> 5))] = v;
            }
        }
        all_idle = __syncthreads_and(idle ? 1 : 0) != 0;
        float s = 0.0f;
        const float* mine = r + g * kW;
        for (int i = 0; i < nw; ++i) s += mine[i];
        phase ^= 1;  // the next reduction writes the other buffer; the one after is separated from these reads by its barrier
        a = s;
    }
    __device__ __forceinline__ void sum1(float& a) { reduce(a, false); }
    __device__ __forceinline__ void sum1_vote(float& a, bool idle) { reduce(a, idle); }
    __device__ __forceinline__ void sum2(float& a, float& b) { reduce(a, false); reduce(b, false); }
    __device__ __forceinline__ bool any(bool p) { float f = p ? 1.0f : 0.0f; reduce(f, false); return f > 0.0f; }
    __device__ __forceinline__ bool all_done(bool) { return all_idle; }
};

template <class F>
__device__ __forceinline__ void stage_params(const SolveArgs& a, int64_t group, float* sp) {
    const int set = a.pset ? a.pset[group] : 0;
    const float* src = a.params + (int64_t)set * F::P;
    F::stage(src, sp, (int)threadIdx.x, (int)blockDim.x);
    __syncthreads();
    if (threadIdx.x == 0) F::prepare(sp);
    __syncthreads();
}

// per-thread parameter-gradient accumulators -> warp shuffle tree -> shared atomics -> one global atomic per
// parameter per CTA
template <class F>
__device__ __forceinline__ void reduce_param_grads(const SolveArgs& a, int64_t group, float* acc, float* sred) {
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < F::P; i += blockDim.x) sred[i] = 0.0f;
    __syncthreads();
    if constexpr (F::kAccInRegs) {
#pragma unroll
        for (int p = 0; p < F::P; ++p) {
            float v = acc[p];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(&sred[p], v);
        }
    } else {
#pragma unroll 1
        for (int p = 0; p < F::P; ++p) {
            float v = acc[p];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(&sred[p], v);
        }
    }
    __syncthreads();
    const int set = a.pset ? a.pset[group] : 0;
    float* dst = a.grad_params + (int64_t)set * F::P;
    for (int i = threadIdx.x; i < F::P; i += blockDim.x) atomicAdd(&dst[i], sred[i]);
}

template <class F>
__device__ __forceinline__ void zero_acc(float* acc) {
    if constexpr (F::kAccInRegs) {
#pragma unroll
        for (int p = 0; p < F::P; ++p) acc[p] = 0.0f;
    } else {
#pragma unroll 1
        for (int p = 0; p < F::P; ++p) acc[p] = 0.0f;
    }
}

template <int ND>
__device__ __forceinline__ DoseReg<ND> load_dose_reg(const SolveArgs& a, int64_t idx) {
    DoseReg<ND> ds;
    ds.amt = a.dose_amt[idx];
#pragma unroll
    for (int j = 0; j < ND; ++j) ds.tau[j] = a.dose_t[idx * a.dose_t_stride + j];
    return ds;
}
__device__ __forceinline__ DoseMem load_dose_mem(const SolveArgs& a, int64_t idx) {
    DoseMem ds;
    ds.amt = a.dose_amt[idx];
    ds.tau = a.dose_t + idx * a.dose_t_stride;
    ds.nd = a.n_dose;
    return ds;
}

// CTA -> (group, tile) decomposition shared by all kernels
struct Tile {
    int64_t group;
    int64_t b;  // trajectory within the group (may be >= batch for padding threads)
};
__device__ __forceinline__ Tile tile_of(const SolveArgs& a, int tiles_per_group) {
    Tile t;
    t.group = blockIdx.x / tiles_per_group;
    t.b = (int64_t)(blockIdx.x % tiles_per_group) * blockDim.x + threadIdx.x;
    return t;
}

// ---------------------------------------------------------------------------------------------------------------
// kernels.  CP: parameters come from the constant bank (ParamConst) instead of the CTA's shared-memory copy.
// ---------------------------------------------------------------------------------------------------------------
#define HODE_WITH_DOSE(ND, a, idx, CALL)                                                   \
    do {                                                                                   \
        if (ND > 0) {                                                                      \
            const DoseReg<(ND > 0 ? ND : 1)> ds = load_dose_reg<(ND > 0 ? ND : 1)>(a, idx); \
            CALL;                                                                          \
        } else {                                                                           \
            const DoseMem ds = load_dose_mem(a, idx);                                      \
            CALL;                                                                          \
        }                                                                                  \
    } while (0)

template <class F, int METHOD, int ND, bool CP>
__global__ void __launch_bounds__(128, HODE_FWD_MINBLOCKS) fixed_fwd_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ float smem[];
    const Tile tl = tile_of(a, tiles_per_group);
    if constexpr (!CP) stage_params<F>(a, tl.group, smem);
    if (tl.b >= a.batch) return;
    const int64_t idx = tl.group * a.batch + tl.b;
    if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (fixed_fwd_traj<F, METHOD>(a, ParamConst(), ds, idx))); }
    else { HODE_WITH_DOSE(ND, a, idx, (fixed_fwd_traj<F, METHOD>(a, (const float*)smem, ds, idx))); }
}

// Forward fixed-grid solve with the read-out + masked SSE consumed at every output time (SseSink, hode_sse.cuh): one launch
// produces loss, grad_h, grad_W, grad_b and the tape; the latent solution is only written if the caller asks for it.
// One parameter set (constant bank), 128-thread CTAs, every lane of every warp runs (padding lanes on a clamped trajectory
// with an empty tile row): the sink is a warp-wide cooperation.
template <class F, int METHOD, int OBS>
__global__ void __launch_bounds__(128, HODE_SSE_MINBLOCKS) fixed_fwd_sse_kernel(const SolveArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int D = F::D;
    using Sink = SseSink<D, OBS>;
    // shared memory: 4 warps x (x tile, mask tile) | sred [OBS * D + OBS] | accumulators [kAccFloats][128] | 4 mbarriers
    constexpr int kSred = 4 * Sink::kTileFloats, kAcc = kSred + round4(OBS * D + OBS), kBar = kAcc + Sink::kAccFloats * 128;
    const int warp = threadIdx.x >> 5;
    float* sred = smem + kSred;
    const int64_t n_traj = a.batch;  // flat launch: n_groups == 1
    const int64_t first = ((int64_t)blockIdx.x * 4 + warp) * 32;
    const int64_t mine = first + (threadIdx.x & 31);
    const int64_t idx = mine < n_traj ? mine : n_traj - 1;
    for (int i = threadIdx.x; i < OBS * D + OBS; i += blockDim.x) sred[i] = 0.0f;
    Sink sink;
    sink.init(a, warp * Sink::kTileFloats, kAcc, kBar + 2 * warp, first, n_traj);
    {
        const DoseReg<1> ds = load_dose_reg<1>(a, idx);
        fixed_fwd_traj<F, METHOD>(a, ParamConst(), ds, idx, sink, /*write_tape=*/mine < n_traj);  // padding lanes write nothing
    }
    __syncthreads();  // sred zeroed by all, solves done
    sink.flush(sred, a.sse_loss, a.sse_inv_norm);
    __syncthreads();
    if (a.sse_grad_w != nullptr) {
        for (int i = threadIdx.x; i < OBS * D; i += blockDim.x) atomicAdd(&a.sse_grad_w[i], sred[i]);
        if (a.sse_grad_b != nullptr)
            for (int i = threadIdx.x; i < OBS; i += blockDim.x) atomicAdd(&a.sse_grad_b[i], sred[OBS * D + i]);
    }
}
template <int D, int OBS>
constexpr size_t sse_smem_bytes() {
    using Sink = SseSink<D, OBS>;
    return (4 * (size_t)Sink::kTileFloats + round4(OBS * D + OBS) + (size_t)Sink::kAccFloats * 128 + 8) * sizeof(float);
}

// Does the field accumulate its parameter gradients warp-cooperatively?  NeuralODE (846 - 3 132 parameters) always;
// RocheODE at D = 12 (104 ml_net accumulators) in the dopri5 reverse sweep and the continuous adjoint, where the
// per-thread accumulators spill (the fixed-grid reverse sweep keeps them: 3 recomputed stages fit next to them).
template <class F> struct CoopOf { static constexpr bool value = false; using type = void; };
template <int D> struct CoopOf<Neural<D>> { static constexpr bool value = true; using type = NeuralCoop<D>; };
template <class F> struct CoopD5 : CoopOf<F> {};
template <bool H, bool A> struct CoopD5<Roche<12, H, A>> { static constexpr bool value = true; using type = RocheCoop<12>; };

template <int D>
__device__ __forceinline__ void coop_init(NeuralCoop<D>& cp, float* stage_base) {
    cp.lane = threadIdx.x & 31;
    cp.stage = stage_base + (threadIdx.x >> 5) * NeuralCoop<D>::kStageFloats;
#pragma unroll
    for (int c = 0; c < NeuralCoop<D>::NJ; ++c) {
#pragma unroll
        for (int i = 0; i < NeuralCoop<D>::SI; ++i) cp.w1[c][i] = 0.0f;
#pragma unroll
        for (int d = 0; d < NeuralCoop<D>::SU; ++d) cp.w2[c][d] = 0.0f;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) cp.b2[d] = 0.0f;
    cp.mute = false;
}
// owned rows -> shared accumulator (one atomic per parameter per warp) -> global
template <int D>
__device__ __forceinline__ void coop_flush(const SolveArgs& a, int64_t group, NeuralCoop<D>& cp, float* sred) {
    using F = Neural<D>;
    for (int i = threadIdx.x; i < F::P; i += blockDim.x) sred[i] = 0.0f;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NeuralCoop<D>::NJ; ++c) {
        const int j = c * 32 + cp.lane;
        if (j < F::H) {
#pragma unroll
            for (int i = 0; i < F::IN; ++i) atomicAdd(&sred[F::OFF_W1 + j * F::IN + i], cp.w1[c][i]);
            atomicAdd(&sred[F::OFF_B1 + j], cp.w1[c][F::IN]);
#pragma unroll
            for (int d = 0; d < D; ++d) atomicAdd(&sred[F::OFF_W2 + d * F::H + j], cp.w2[c][d]);
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float v = cp.b2[d];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (cp.lane == 0) atomicAdd(&sred[F::OFF_B2 + d], v);
    }
    __syncthreads();
    const int set = a.pset ? a.pset[group] : 0;
    float* dst = a.grad_params + (int64_t)set * F::P;
    for (int i = threadIdx.x; i < F::P; i += blockDim.x) atomicAdd(&dst[i], sred[i]);
}

template <int D>
__device__ __forceinline__ void coop_init(RocheCoop<D>& cp, float* stage_base) {
    cp.lane = threadIdx.x & 31;
    cp.stage = stage_base + (threadIdx.x >> 5) * RocheCoop<D>::kStageFloats;
#pragma unroll
    for (int c = 0; c < 6; ++c) { cp.w[c][0] = 0.0f; cp.w[c][1] = 0.0f; }
    cp.b[0] = 0.0f; cp.b[1] = 0.0f;
#pragma unroll
    for (int i = 0; i < RocheCoop<D>::NE; ++i) cp.e[i] = 0.0f;
    cp.mute = false;
}
// owned blocks (4 partial sums per entry, one per trajectory group of the warp) + per-thread expert scalars -> shared
// accumulator -> one global atomic per parameter per CTA
template <class F, bool EG, int D>
__device__ __forceinline__ void coop_flush(const SolveArgs& a, int64_t group, RocheCoop<D>& cp, float* sred) {
    for (int i = threadIdx.x; i < F::P; i += blockDim.x) sred[i] = 0.0f;
    __syncthreads();
    const int jb = (cp.lane & 7) >> 1, db = cp.lane & 1;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        atomicAdd(&sred[F::OFF_W + (2 * jb) * D + 6 * db + c], cp.w[c][0]);
        atomicAdd(&sred[F::OFF_W + (2 * jb + 1) * D + 6 * db + c], cp.w[c][1]);
    }
    if (db == 0) {
        atomicAdd(&sred[F::OFF_B + 2 * jb], cp.b[0]);
        atomicAdd(&sred[F::OFF_B + 2 * jb + 1], cp.b[1]);
    }
    if (EG) {
#pragma unroll
        for (int i = 0; i < RocheCoop<D>::NE; ++i) {
            const int pi = i < R_NSCALAR ? i : F::OFF_TH + (i - R_NSCALAR);  // 13 expert scalars, then theta_1 / theta_2
            if (pi < F::P && (i < R_NSCALAR || F::ABLATE) && i < R_NSCALAR + 2) {
                float v = cp.e[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (cp.lane == 0) atomicAdd(&sred[pi], v);
            }
        }
    }
    __syncthreads();
    const int set = a.pset ? a.pset[group] : 0;
    float* dst = a.grad_params + (int64_t)set * F::P;
    for (int i = threadIdx.x; i < F::P; i += blockDim.x) atomicAdd(&dst[i], sred[i]);
}

// CTAs of 128 threads per SM that the reverse sweeps' register budget must allow.  Left alone, ptxas takes 215 registers for
// the NeuralODE sweep at D = 6 (2 CTAs per SM, issue-active 41 %, `wait` 2.0 per issue); capped at 128 registers it spills
// 24 bytes and runs 1.29 x faster (34.7 -> 26.9 ms at 131 072 patients).
template <class F> struct BwdBlocks { static constexpr int value = HODE_BWD_MINBLOCKS; };
template <int D> struct BwdBlocks<Neural<D>> { static constexpr int value = D <= 6 ? 4 : (D <= 8 ? 3 : HODE_BWD_MINBLOCKS); };

template <class F, int METHOD, bool EG, int ND, bool CP>
__global__ void __launch_bounds__(128, BwdBlocks<F>::value) fixed_bwd_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ __align__(16) float smem[];
    float* sp = smem;
    float* sred = smem + (CP ? 0 : round4(F::SP));
    const Tile tl = tile_of(a, tiles_per_group);
    if constexpr (!CP) stage_params<F>(a, tl.group, sp);
    if constexpr (CoopD5<F>::value) {
        // NeuralODE, and RocheODE at D = 12 (104 ml_net accumulators: 255 registers + spills when kept per thread): every lane
        // of every warp runs the sweep (padding lanes on a clamped trajectory with zero gradients) -- the parameter-gradient
        // accumulation is a warp-wide cooperation
        typename CoopD5<F>::type cp;
        coop_init(cp, sred + round4(F::P));
        const bool valid = tl.b < a.batch;
        const int64_t idx = tl.group * a.batch + (valid ? tl.b : a.batch - 1);
        if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (fixed_bwd_traj<F, METHOD, EG>(a, ParamConst(), ds, idx, &cp, valid))); }
        else { HODE_WITH_DOSE(ND, a, idx, (fixed_bwd_traj<F, METHOD, EG>(a, (const float*)sp, ds, idx, &cp, valid))); }
        if constexpr (CoopOf<F>::value) coop_flush(a, tl.group, cp, sred);
        else coop_flush<F, EG>(a, tl.group, cp, sred);
    } else {
        float acc[F::P];
        zero_acc<F>(acc);
        if (tl.b < a.batch) {
            const int64_t idx = tl.group * a.batch + tl.b;
            if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (fixed_bwd_traj<F, METHOD, EG>(a, ParamConst(), ds, idx, (float*)acc))); }
            else { HODE_WITH_DOSE(ND, a, idx, (fixed_bwd_traj<F, METHOD, EG>(a, (const float*)sp, ds, idx, (float*)acc))); }
        }
        reduce_param_grads<F>(a, tl.group, acc, sred);
    }
}

// continuous adjoint of a fixed-grid solve: same shape as the reverse sweep, no tape (fixed_adj_traj)
template <class F, int METHOD, bool EG, int ND, bool CP>
__global__ void __launch_bounds__(128, BwdBlocks<F>::value) fixed_adj_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ __align__(16) float smem[];
    float* sp = smem;
    float* sred = smem + (CP ? 0 : round4(F::SP));
    const Tile tl = tile_of(a, tiles_per_group);
    if constexpr (!CP) stage_params<F>(a, tl.group, sp);
    if constexpr (CoopD5<F>::value) {
        // NeuralODE, and RocheODE at D = 12 (104 ml_net accumulators: 255 registers + spills when kept per thread): every lane
        // of every warp runs the sweep (padding lanes on a clamped trajectory with zero gradients) -- the parameter-gradient
        // accumulation is a warp-wide cooperation
        typename CoopD5<F>::type cp;
        coop_init(cp, sred + round4(F::P));
        const bool valid = tl.b < a.batch;
        const int64_t idx = tl.group * a.batch + (valid ? tl.b : a.batch - 1);
        if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (fixed_adj_traj<F, METHOD, EG>(a, ParamConst(), ds, idx, &cp, valid))); }
        else { HODE_WITH_DOSE(ND, a, idx, (fixed_adj_traj<F, METHOD, EG>(a, (const float*)sp, ds, idx, &cp, valid))); }
        if constexpr (CoopOf<F>::value) coop_flush(a, tl.group, cp, sred);
        else coop_flush<F, EG>(a, tl.group, cp, sred);
    } else {
        float acc[F::P];
        zero_acc<F>(acc);
        if (tl.b < a.batch) {
            const int64_t idx = tl.group * a.batch + tl.b;
            if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (fixed_adj_traj<F, METHOD, EG>(a, ParamConst(), ds, idx, (float*)acc))); }
            else { HODE_WITH_DOSE(ND, a, idx, (fixed_adj_traj<F, METHOD, EG>(a, (const float*)sp, ds, idx, (float*)acc))); }
        }
        reduce_param_grads<F>(a, tl.group, acc, sred);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dopri5 kernels.  Where do the stage rows live?  In registers for D <= 8 (stage loops unrolled); in shared memory for
// D = 12 (RowsMem, stage loops rolled): 7 x 12 stage floats on top of state + stage input + error estimate put the
// register-resident forward kernel at 240 registers (2 warps per scheduler, ncu: issue-active 42 %, `wait` 1.7 per issue)
// and the reverse sweep at 255 registers + 1.9 KB of spills with 118 KB of unrolled code (stall_no_instruction 2.1).
// ---------------------------------------------------------------------------------------------------------------
template <class F>
struct D5Store {
    static constexpr bool kSmem = F::D >= 12;
    // Rolled stage loops wherever the rows are in shared memory: one inlined copy of the vector field (and of its VJP)
    // per loop, stage algebra dispatched to per-stage straight-line code (stage_switch).  Unrolling the forward loop with
    // shared-memory rows was tried: ptxas hoists the 34 tableau x dt products and the row loads across stages and needs
    // ~250 registers again (or spills at the 128 / 168 caps).
    static constexpr bool kFwdRolled = kSmem, kBwdRolled = kSmem;
    static constexpr int kFwdRows = 7, kBwdRows = 9;
    static constexpr int kFwdMinBlocks = kSmem ? HODE_D5_FWD_MINBLOCKS : (F::D <= 6 ? HODE_D5_FWD_MINBLOCKS_REG : 1);  // 128-thread CTAs per SM the register budget must allow
    static constexpr int kBwdMinBlocks = kSmem ? 3 : 1;
    __host__ __device__ static constexpr size_t fwd_floats(int threads) { return kSmem ? (size_t)kFwdRows * F::D * threads : 0; }
    __host__ __device__ static constexpr size_t bwd_floats(int threads) { return kSmem ? (size_t)kBwdRows * F::D * threads : 0; }
};

// runs `fn(rows)` with the storage D5Store selects; `base` = the CTA's row area in shared memory (16-byte aligned)
// THREADS > 0: the CTA size is known at compile time (row strides become immediates)
template <class F, int NR, int THREADS, class Fn>
__device__ __forceinline__ void with_rows(float* base, Fn&& fn) {
    if constexpr (D5Store<F>::kSmem) {
        using RM = RowsMem<F::D, NR, THREADS * RowsMem<F::D, NR>::VEC>;
        RM rows{base + threadIdx.x * RM::VEC, (int)blockDim.x * RM::VEC};
        fn(rows);
    } else {
        RowsReg<F::D, NR> rows;
        fn(rows);
    }
}

// MAXT: launch bound.  128 for per-trajectory control and small groups; HODE_DOPRI5_MAX_THREADS for a batch-coupled group
// that needs a whole large CTA (128 registers per thread).
template <class F, bool PER_TRAJ, int ND, int MAXT, bool CP>
__global__ void __launch_bounds__(MAXT, MAXT <= 128 ? D5Store<F>::kFwdMinBlocks : 1)
dopri5_fwd_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ __align__(16) float smem[];
    float* sp = smem;
    float* red = smem + (CP ? 0 : round4(F::SP));
    float* rows = red + 128;
    const Tile tl = tile_of(a, tiles_per_group);
    if constexpr (!CP) stage_params<F>(a, tl.group, sp);
    const bool valid = tl.b < a.batch;
    if (PER_TRAJ && !valid) return;
    const int64_t b = valid ? tl.b : (a.batch - 1);
    const int64_t idx = tl.group * a.batch + b;
    const int64_t ctrl = PER_TRAJ ? idx : tl.group;
    const bool leader = PER_TRAJ ? true : (threadIdx.x == 0);
    const float count = PER_TRAJ ? (float)F::D : (float)(a.batch * F::D);
    with_rows<F, 7, (MAXT <= 128 ? 128 : 0)>(rows, [&](auto& k) {
        if (PER_TRAJ) {
            CommNone cm;
            if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, ParamConst(), ds, k, idx, valid, ctrl, leader, count))); }
            else { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, (const float*)sp, ds, k, idx, valid, ctrl, leader, count))); }
        } else {
            CommCta cm{red, (int)(blockDim.x >> 5), 0};
            if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, ParamConst(), ds, k, idx, valid, ctrl, leader, count))); }
            else { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, (const float*)sp, ds, k, idx, valid, ctrl, leader, count))); }
        }
    });
}

// Batch-coupled controller for SMALL groups (batch <= 16, one parameter set): floor(32 / batch) groups per warp, each a
// lane segment with its own step-size sequence.  (One group per CTA leaves 22 of 32 lanes idle at the reference's
// batch of 10, run_dim.sh:41.)
template <class F, int ND, bool CP>
__global__ void __launch_bounds__(128, D5Store<F>::kFwdMinBlocks) dopri5_fwd_seg_kernel(const SolveArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* segbuf = smem + (CP ? 0 : round4(F::SP));
    float* rows = segbuf + 4 * CommSeg::kFloatsPerWarp;
    if constexpr (!CP) stage_params<F>(a, 0, smem);
    for (int i = threadIdx.x; i < 4 * CommSeg::kFloatsPerWarp; i += blockDim.x) segbuf[i] = 0.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int size = (int)a.batch, gpw = 32 / size, seg = lane / size;
    if (seg >= gpw) return;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t group = warp * gpw + seg;
    if (group >= a.n_groups) return;
    CommSeg cm;
    cm.part = __activemask();  // lanes left after the early returns above (converged here)
    const int base = seg * size;
    cm.rel = lane - base;
    cm.nvec = (size + 3) >> 2;
    cm.mask = (size == 32 ? 0xffffffffu : ((1u << size) - 1u)) << base;
    cm.buf = segbuf + (threadIdx.x >> 5) * CommSeg::kFloatsPerWarp + seg * 32;
    const int64_t idx = group * a.batch + cm.rel;
    const float count = (float)(a.batch * F::D);
    with_rows<F, 7, 128>(rows, [&](auto& k) {
        if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, ParamConst(), ds, k, idx, true, group, cm.rel == 0, count))); }
        else { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, (const float*)smem, ds, k, idx, true, group, cm.rel == 0, count))); }
    });
}

// Batch-coupled controller for MID-SIZE groups (17 <= batch <= 128, one parameter set) whose size wastes lanes as one CTA
// per group: floor(256 / batch) groups per 256-thread CTA (CommPack).
template <class F, int ND, bool CP>
__global__ void __launch_bounds__(CommPack::kThreads, D5Store<F>::kSmem || F::D <= 6 ? 2 : 1) dopri5_fwd_pack_kernel(const SolveArgs a, int gpc) {
    extern __shared__ __align__(16) float smem[];
    float* red = smem + (CP ? 0 : round4(F::SP));
    float* rows = red + CommPack::kFloats;
    if constexpr (!CP) stage_params<F>(a, 0, smem);
    const int tid = threadIdx.x, size = (int)a.batch;
    const int gl = tid / size;  // group inside the CTA
    const int64_t group = (int64_t)blockIdx.x * gpc + gl;
    const bool valid = gl < gpc && group < a.n_groups;
    const int rel = tid - gl * size;
    CommPack cm;
    cm.red = red;
    cm.batch = size;
    cm.g = valid ? gl : CommPack::kG - 1;  // gpc <= 15: slot kG - 1 is never written
    const int w0 = (tid >> 5) << 5;
    cm.g_lo = w0 / size;
    cm.g_hi = min((w0 + 31) / size, gpc - 1);
    cm.nw = valid ? (((gl * size + size - 1) >> 5) - ((gl * size) >> 5) + 1) : 0;
    const int64_t idx = valid ? group * a.batch + rel : 0;
    const float count = (float)(a.batch * F::D);
    with_rows<F, 7, 0>(rows, [&](auto& k) {
        if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, ParamConst(), ds, k, idx, valid, valid ? group : 0, valid && rel == 0, count))); }
        else { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, (const float*)smem, ds, k, idx, valid, valid ? group : 0, valid && rel == 0, count))); }
    });
}

// Batch-coupled controller for a group of more than 512 trajectories: `nrank` CTAs of one cluster per group (CommCluster).
template <class F, int ND, bool CP>
__global__ void __launch_bounds__(HODE_DOPRI5_MAX_THREADS, 1) dopri5_fwd_cluster_kernel(const SolveArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* red = smem + (CP ? 0 : round4(F::SP));
    float* slots = red + 128;
    float* rows = slots + CommCluster::kFloats;
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned nrank = cl.num_blocks(), rank = cl.block_rank();
    const int64_t group = blockIdx.x / nrank;
    if constexpr (!CP) stage_params<F>(a, group, smem);
    const int64_t b0 = (int64_t)rank * blockDim.x + threadIdx.x;
    const bool valid = b0 < a.batch;
    const int64_t idx = group * a.batch + (valid ? b0 : a.batch - 1);
    const float count = (float)(a.batch * F::D);
    CommCluster cm{CommCta{red, (int)(blockDim.x >> 5), 0}, slots, rank, nrank, 0};
    with_rows<F, 7, 0>(rows, [&](auto& k) {
        if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, ParamConst(), ds, k, idx, valid, group, rank == 0 && threadIdx.x == 0, count))); }
        else { HODE_WITH_DOSE(ND, a, idx, (dopri5_fwd_traj<F, D5Store<F>::kFwdRolled>(a, cm, (const float*)smem, ds, k, idx, valid, group, rank == 0 && threadIdx.x == 0, count))); }
    });
    cl.sync();  // no CTA leaves while a sibling could still address its shared memory
}

template <class F, bool EG, int ND, bool CP>
__global__ void __launch_bounds__(128, D5Store<F>::kBwdMinBlocks) dopri5_bwd_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ __align__(16) float smem[];
    float* sp = smem;
    float* sred = smem + (CP ? 0 : round4(F::SP));
    float* rows = sred + round4(F::P);
    float* coop_stage = rows + D5Store<F>::bwd_floats(128);
    const Tile tl = tile_of(a, tiles_per_group);
    if constexpr (!CP) stage_params<F>(a, tl.group, sp);
    const bool valid = tl.b < a.batch;
    const int64_t idx = tl.group * a.batch + (valid ? tl.b : a.batch - 1);
    const int64_t ctrl = a.per_traj ? idx : (a.ctrl_batch > 0 ? idx / a.ctrl_batch : tl.group);
    const int my_steps = valid ? a.stats[ctrl].accepted : 0;
    if constexpr (CoopD5<F>::value) {
        // Cooperative accumulators: every lane of every warp walks the warp's longest tape (padding lanes and lanes with
        // fewer accepted steps idle with zero adjoints).  Neural: batch-coupled groups only (one group per CTA).
        constexpr bool kRoche = !CoopOf<F>::value;
        if (kRoche || !a.per_traj) {
            typename CoopD5<F>::type cp;
            coop_init(cp, coop_stage);
            const int nloop = __reduce_max_sync(0xffffffffu, my_steps);
            with_rows<F, 9, 128>(rows, [&](auto& R) {
                if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (dopri5_bwd_traj<F, EG, D5Store<F>::kBwdRolled>(a, ParamConst(), ds, R, idx, ctrl, &cp, valid, nloop))); }
                else { HODE_WITH_DOSE(ND, a, idx, (dopri5_bwd_traj<F, EG, D5Store<F>::kBwdRolled>(a, (const float*)sp, ds, R, idx, ctrl, &cp, valid, nloop))); }
            });
            if constexpr (kRoche) coop_flush<F, EG>(a, tl.group, cp, sred);
            else coop_flush(a, tl.group, cp, sred);
            return;
        }
    }
    float acc[F::P];
    zero_acc<F>(acc);
    if (valid) {
        with_rows<F, 9, 128>(rows, [&](auto& R) {
            if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (dopri5_bwd_traj<F, EG, D5Store<F>::kBwdRolled>(a, ParamConst(), ds, R, idx, ctrl, (float*)acc, true, my_steps))); }
            else { HODE_WITH_DOSE(ND, a, idx, (dopri5_bwd_traj<F, EG, D5Store<F>::kBwdRolled>(a, (const float*)sp, ds, R, idx, ctrl, (float*)acc, true, my_steps))); }
        });
    }
    reduce_param_grads<F>(a, tl.group, acc, sred);
}

// Adaptive continuous adjoint (dopri5_adj_traj): one controller per CTA (batch-coupled, torchdiffeq semantics) or per
// trajectory (flat launch).  14 rows per thread always live in shared memory (7 field values + 7 adjoint derivatives).
// Parameter gradients: warp-cooperative accumulators where the field has them and the warp is converged (batch-coupled
// groups: accept / reject is CTA-uniform); per-thread accumulators otherwise.
// fields for which the mixed-norm adaptive adjoint is built: per-thread parameter accumulators (RocheODE up to D = 8)
template <class F> struct MixedOk { static constexpr bool value = F::kAccInRegs && !CoopD5<F>::value; };
template <class F>
constexpr size_t mixed_floats(int threads) {
    return (size_t)round4(F::P) + round4(2 * F::P) + 4 + (size_t)(threads / 32) * 2 * F::P;
}

template <class F, bool PER_TRAJ, bool EG, int ND, int MAXT, bool CP>
__global__ void __launch_bounds__(MAXT, MAXT <= 128 ? 2 : 1) dopri5_adj_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ __align__(16) float smem[];
    float* sp = smem;
    float* red = smem + (CP ? 0 : round4(F::SP));
    float* sred = red + 128;
    float* rows = sred + round4(F::P);
    using RM = RowsMem<F::D, 14>;
    float* coop_stage = rows + (size_t)RM::kFloatsPerThread * blockDim.x;
    const Tile tl = tile_of(a, tiles_per_group);
    if constexpr (!CP) stage_params<F>(a, tl.group, sp);
    const bool valid = tl.b < a.batch;
    const int64_t b = valid ? tl.b : (a.batch - 1);
    const int64_t idx = tl.group * a.batch + b;
    const int64_t ctrl = PER_TRAJ ? idx : tl.group;
    const bool leader = PER_TRAJ ? valid : (threadIdx.x == 0);
    const float count = PER_TRAJ ? (float)F::D : (float)(a.batch * F::D);
    RM R{rows + threadIdx.x * RM::VEC, (int)blockDim.x * RM::VEC};
    auto run = [&](auto& cm, auto accp) {
        if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (dopri5_adj_traj<F, EG, true>(a, cm, ParamConst(), ds, R, idx, valid, ctrl, leader, count, accp))); }
        else { HODE_WITH_DOSE(ND, a, idx, (dopri5_adj_traj<F, EG, true>(a, cm, (const float*)sp, ds, R, idx, valid, ctrl, leader, count, accp))); }
    };
    if constexpr (MixedOk<F>::value && !PER_TRAJ) {
        if (a.adj_mixed) {  // torchdiffeq's default adjoint norm: the group's parameter adjoint lives in shared memory
            float* pg = coop_stage;  // [P] g, [2 P] reduction results, [4] scalars, then the per-warp partials of sum_vec
            ParamCtl pc{pg, pg + round4(F::P), pg + round4(F::P) + round4(2 * F::P)};
            CommCta cm{red, (int)(blockDim.x >> 5), 0};
            cm.vbuf = pc.r + 4;
            if constexpr (CP) { HODE_WITH_DOSE(ND, a, idx, (dopri5_adj_mixed_traj<F, EG>(a, cm, ParamConst(), ds, R, idx, valid, ctrl, leader, count, pc, (int)threadIdx.x, (int)blockDim.x))); }
            else { HODE_WITH_DOSE(ND, a, idx, (dopri5_adj_mixed_traj<F, EG>(a, cm, (const float*)sp, ds, R, idx, valid, ctrl, leader, count, pc, (int)threadIdx.x, (int)blockDim.x))); }
            __syncthreads();
            const int set = a.pset ? a.pset[tl.group] : 0;
            for (int k = threadIdx.x; k < F::P; k += blockDim.x) atomicAdd(a.grad_params + (int64_t)set * F::P + k, pc.g[k]);
            return;
        }
    }
    if constexpr (!PER_TRAJ && CoopD5<F>::value) {
        typename CoopD5<F>::type cp;
        coop_init(cp, coop_stage);
        CommCta cm{red, (int)(blockDim.x >> 5), 0};
        run(cm, &cp);
        if constexpr (!CoopOf<F>::value) coop_flush<F, EG>(a, tl.group, cp, sred);
        else coop_flush(a, tl.group, cp, sred);
    } else {
        float acc[F::P];
        zero_acc<F>(acc);
        if (PER_TRAJ) {
            CommNone cm;
            if (valid) run(cm, (float*)acc);
        } else {
            CommCta cm{red, (int)(blockDim.x >> 5), 0};
            run(cm, (float*)acc);
        }
        if (!valid) zero_acc<F>(acc);
        reduce_param_grads<F>(a, tl.group, acc, sred);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// launchers (explicitly instantiated per field in inst_*.cu)
// ---------------------------------------------------------------------------------------------------------------
inline int round_up32(int64_t n) { return (int)(((n + 31) / 32) * 32); }
template <class K>
inline int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return 0;
}

#define HODE_LAUNCH_CHECK()                                   \
    do {                                                      \
        cudaError_t e_ = cudaGetLastError();                  \
        if (e_ != cudaSuccess) return (int)e_;                \
    } while (0)

// The n_dose == 1 specialisation (dose schedule in registers) doubles the number of kernels of a field.  It is compiled for
// the variants every reference call site uses (Hill exponents of 2, the NeuralODE field); the generic-Hill-exponent and ablation
// variants of RocheODE -- taken only when somebody trains the Hill exponents / runs the ablation study -- use the general
// dose path and shared-memory parameters (Roche::kConstBank), which quarters their share of the build.
template <class F> struct NdSpecial { static constexpr bool value = true; };
template <int D, bool A> struct NdSpecial<Roche<D, false, A>> { static constexpr bool value = false; };
template <int D> struct NdSpecial<Roche<D, true, true>> { static constexpr bool value = false; };
#define HODE_ND(A, B)                                 \
    do {                                              \
        if constexpr (NdSpecial<F>::value) {          \
            if (nd1) { A; } else { B; }               \
        } else {                                      \
            B;                                        \
        }                                             \
    } while (0)

// Runs `LAUNCH(CPFLAG)` with CPFLAG = true behind a constant-bank lease when the launch qualifies, else CPFLAG = false.
#define HODE_DISPATCH_CP(F, a, st, LAUNCH)                               \
    do {                                                                 \
        if constexpr (F::kConstBank) {                                   \
            if (use_const_params<F>(a)) {                                \
                ConstParamLease lease((F*)nullptr, a.params, st);        \
                if (lease.rc != 0) return lease.rc;                      \
                LAUNCH(true);                                            \
                lease.done();                                            \
                break;                                                   \
            }                                                            \
        }                                                                \
        LAUNCH(false);                                                   \
    } while (0)

template <class F>
constexpr size_t coop_stage_floats() {
    if constexpr (CoopD5<F>::value) return CoopD5<F>::type::kStageFloats;
    else return 0;
}

// With ONE parameter set nothing ties a CTA to a group: threads enumerate all trajectories ("flat"), so small groups do
// not leave lanes idle.  The trajectory index is unchanged (group * batch + b).
inline SolveArgs flatten(const SolveArgs& a) {
    SolveArgs f = a;
    if (a.pset == nullptr && a.n_groups > 1) {
        f.ctrl_batch = a.batch;
        f.batch = a.n_groups * a.batch;
        f.n_groups = 1;
    }
    return f;
}

template <class F>
int launch_fixed_fwd(const hode_cfg& cfg, const SolveArgs& a_in, cudaStream_t st) {
    const SolveArgs a = flatten(a_in);
    const int threads = a.batch >= 128 ? 128 : round_up32(a.batch);
    const int tiles = (int)((a.batch + threads - 1) / threads);
    const int64_t nblk = a.n_groups * tiles;
    const bool nd1 = cfg.n_dose == 1;
#define HODE_FF(M, ND, CP) fixed_fwd_kernel<F, M, ND, CP><<<(unsigned)nblk, threads, (CP ? 0 : F::SP) * sizeof(float), st>>>(a, tiles)
#define HODE_FF_CP(CP)                                                                                   \
    switch (cfg.method) {                                                                                \
        case HODE_EULER: HODE_ND(HODE_FF(M_EULER, 1, CP), HODE_FF(M_EULER, 0, CP)); break;          \
        case HODE_MIDPOINT: HODE_ND(HODE_FF(M_MIDPOINT, 1, CP), HODE_FF(M_MIDPOINT, 0, CP)); break; \
        case HODE_RK4_38: HODE_ND(HODE_FF(M_RK4_38, 1, CP), HODE_FF(M_RK4_38, 0, CP)); break;       \
        default: return -1;                                                                              \
    }
    HODE_DISPATCH_CP(F, a, st, HODE_FF_CP);
#undef HODE_FF_CP
#undef HODE_FF
    HODE_LAUNCH_CHECK();
    return 0;
}

// -1: this (field, method, obs, n_dose, parameter-set) combination has no fused kernel (the caller falls back to the
// separate solve + decode launches); otherwise a CUDA error code or 0.
template <class F> struct SseOk { static constexpr bool value = false; };
template <int D> struct SseOk<Roche<D, true, false>> { static constexpr bool value = (D == 4 || D == 6 || D == 8); };

template <class F>
int launch_fixed_fwd_sse(const hode_cfg& cfg, const SolveArgs& a_in, cudaStream_t st) {
    if constexpr (!SseOk<F>::value) {
        return -1;
    } else {
        constexpr int D = F::D;
        const SolveArgs a = flatten(a_in);
        const int obs = a.sse_obs;
        if (a.pset != nullptr || a.n_groups != 1 || cfg.n_dose != 1 || !use_const_params<F>(a)) return -1;
        if (obs != 20 && obs != 24 && obs != 40 && obs != 80) return -1;
        if (cfg.method != HODE_EULER && cfg.method != HODE_MIDPOINT && cfg.method != HODE_RK4_38) return -1;
        const int64_t nblk = (a.batch + 127) / 128;
        ConstParamLease lease((F*)nullptr, a.params, st);
        if (lease.rc != 0) return lease.rc;
        prep_readout_kernel<<<1, 128, 0, st>>>(a.sse_w, a.sse_b, obs, D);
        void* stage_ptr = nullptr;
        cudaError_t e = cudaGetSymbolAddress(&stage_ptr, g_readout_stage);
        if (e != cudaSuccess) return (int)e;
        e = cudaMemcpyToSymbolAsync(c_readout, stage_ptr, sizeof(float) * (2 * obs * D + obs), 0, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
#define HODE_FS(M, OBS)                                                                        \
    do {                                                                                       \
        int e_ = set_smem(fixed_fwd_sse_kernel<F, M, OBS>, sse_smem_bytes<D, OBS>());          \
        if (e_ != 0) return e_;                                                                \
        fixed_fwd_sse_kernel<F, M, OBS><<<(unsigned)nblk, 128, sse_smem_bytes<D, OBS>(), st>>>(a); \
    } while (0)
#define HODE_FS_M(M)                                        \
    do {                                                    \
        if (obs == 20) HODE_FS(M, 20);                      \
        else if (obs == 24) HODE_FS(M, 24);                 \
        else if (obs == 40) HODE_FS(M, 40);                 \
        else HODE_FS(M, 80);                                \
    } while (0)
        switch (cfg.method) {
            case HODE_EULER: HODE_FS_M(M_EULER); break;
            case HODE_MIDPOINT: HODE_FS_M(M_MIDPOINT); break;
            default: HODE_FS_M(M_RK4_38); break;
        }
#undef HODE_FS_M
#undef HODE_FS
        lease.done();
        HODE_LAUNCH_CHECK();
        return 0;
    }
}

template <class F>
int launch_fixed_bwd(const hode_cfg& cfg, const SolveArgs& a_in, cudaStream_t st) {
    const SolveArgs a = flatten(a_in);
    const int threads = a.batch >= 128 ? 128 : round_up32(a.batch);
    const int tiles = (int)((a.batch + threads - 1) / threads);
    const int64_t nblk = a.n_groups * tiles;
    const bool nd1 = cfg.n_dose == 1;
    const bool eg = cfg.expert_grads != 0;
#define HODE_FB(M, EG, ND, CP)                                                                                        \
    do {                                                                                                           \
        const size_t sh_ = ((CP ? 0 : round4(F::SP)) + round4(F::P) + coop_stage_floats<F>() * (size_t)(threads / 32)) * sizeof(float); \
        if (sh_ > 48 * 1024)                                                                                       \
            cudaFuncSetAttribute(fixed_bwd_kernel<F, M, EG, ND, CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_); \
        fixed_bwd_kernel<F, M, EG, ND, CP><<<(unsigned)nblk, threads, sh_, st>>>(a, tiles);                        \
    } while (0)
#define HODE_FB_M(M, CP)                                                           \
    do {                                                                           \
        if (eg) { HODE_ND(HODE_FB(M, true, 1, CP), HODE_FB(M, true, 0, CP)); } \
        else    { HODE_ND(HODE_FB(M, false, 1, CP), HODE_FB(M, false, 0, CP)); } \
    } while (0)
#define HODE_FB_CP(CP)                                      \
    switch (cfg.method) {                                   \
        case HODE_EULER: HODE_FB_M(M_EULER, CP); break;     \
        case HODE_MIDPOINT: HODE_FB_M(M_MIDPOINT, CP); break; \
        case HODE_RK4_38: HODE_FB_M(M_RK4_38, CP); break;   \
        default: return -1;                                 \
    }
    HODE_DISPATCH_CP(F, a, st, HODE_FB_CP);
#undef HODE_FB_CP
#undef HODE_FB_M
#undef HODE_FB
    HODE_LAUNCH_CHECK();
    return 0;
}

template <class F>
int launch_fixed_adj(const hode_cfg& cfg, const SolveArgs& a_in, cudaStream_t st) {
    const SolveArgs a = flatten(a_in);
    const int threads = a.batch >= 128 ? 128 : round_up32(a.batch);
    const int tiles = (int)((a.batch + threads - 1) / threads);
    const int64_t nblk = a.n_groups * tiles;
    const bool nd1 = cfg.n_dose == 1;
    const bool eg = cfg.expert_grads != 0;
#define HODE_FA(M, EG, ND, CP)                                                                                        \
    do {                                                                                                           \
        const size_t sh_ = ((CP ? 0 : round4(F::SP)) + round4(F::P) + coop_stage_floats<F>() * (size_t)(threads / 32)) * sizeof(float); \
        if (sh_ > 48 * 1024)                                                                                       \
            cudaFuncSetAttribute(fixed_adj_kernel<F, M, EG, ND, CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_); \
        fixed_adj_kernel<F, M, EG, ND, CP><<<(unsigned)nblk, threads, sh_, st>>>(a, tiles);                        \
    } while (0)
#define HODE_FA_M(M, CP)                                                           \
    do {                                                                           \
        if (eg) { HODE_ND(HODE_FA(M, true, 1, CP), HODE_FA(M, true, 0, CP)); } \
        else    { HODE_ND(HODE_FA(M, false, 1, CP), HODE_FA(M, false, 0, CP)); } \
    } while (0)
#define HODE_FA_CP(CP)                                      \
    switch (cfg.method) {                                   \
        case HODE_EULER: HODE_FA_M(M_EULER, CP); break;     \
        case HODE_MIDPOINT: HODE_FA_M(M_MIDPOINT, CP); break; \
        case HODE_RK4_38: HODE_FA_M(M_RK4_38, CP); break;   \
        default: return -1;                                 \
    }
    HODE_DISPATCH_CP(F, a, st, HODE_FA_CP);
#undef HODE_FA_CP
#undef HODE_FA_M
#undef HODE_FA
    HODE_LAUNCH_CHECK();
    return 0;
}

template <class F>
int launch_dopri5_fwd(const hode_cfg& cfg, const SolveArgs& a_in, cudaStream_t st) {
    const bool nd1 = cfg.n_dose == 1;
    if (!a_in.per_traj && a_in.batch > (int64_t)HODE_DOPRI5_MAX_THREADS * CommCluster::kMaxCtas) return -2;
    if (!a_in.per_traj && a_in.batch > HODE_DOPRI5_MAX_THREADS) {
        // one cluster of CTAs per group
        const SolveArgs& a = a_in;
        const int nrank = (int)((a.batch + HODE_DOPRI5_MAX_THREADS - 1) / HODE_DOPRI5_MAX_THREADS);
        const int threads = round_up32((int)((a.batch + nrank - 1) / nrank));
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)(a.n_groups * nrank), 1, 1);
        lc.blockDim = dim3((unsigned)threads, 1, 1);
        lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)nrank; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
#define HODE_DC(ND, CP)                                                                                                   \
    do {                                                                                                                  \
        const size_t sh_ = ((CP ? 0 : round4(F::SP)) + 128 + CommCluster::kFloats + D5Store<F>::fwd_floats(threads)) * sizeof(float); \
        int e_ = set_smem(dopri5_fwd_cluster_kernel<F, ND, CP>, sh_);                                                     \
        if (e_ != 0) return e_;                                                                                           \
        lc.dynamicSmemBytes = sh_;                                                                                        \
        cudaError_t le_ = cudaLaunchKernelEx(&lc, dopri5_fwd_cluster_kernel<F, ND, CP>, a);                               \
        if (le_ != cudaSuccess) return (int)le_;                                                                          \
    } while (0)
#define HODE_DC_CP(CP) do { HODE_ND(HODE_DC(1, CP), HODE_DC(0, CP)); } while (0)
        HODE_DISPATCH_CP(F, a, st, HODE_DC_CP);
#undef HODE_DC_CP
#undef HODE_DC
        HODE_LAUNCH_CHECK();
        return 0;
    }
    if (!a_in.per_traj && a_in.pset == nullptr && a_in.batch <= 16 && a_in.batch >= 4) {
        // small batch-coupled groups: several groups per warp
        const SolveArgs& a = a_in;
        const int gpw = 32 / (int)a.batch;
        const int64_t warps = (a.n_groups + gpw - 1) / gpw;
        const int threads = warps >= 4 ? 128 : (int)warps * 32;
        const int64_t nblk = (warps * 32 + threads - 1) / threads;
#define HODE_DS(ND, CP)                                                                                                          \
    do {                                                                                                                         \
        const size_t sh_ = ((CP ? 0 : round4(F::SP)) + 4 * CommSeg::kFloatsPerWarp + D5Store<F>::fwd_floats(128)) * sizeof(float); \
        int e_ = set_smem(dopri5_fwd_seg_kernel<F, ND, CP>, sh_);                                                                \
        if (e_ != 0) return e_;                                                                                                  \
        dopri5_fwd_seg_kernel<F, ND, CP><<<(unsigned)nblk, threads, sh_, st>>>(a);                                               \
    } while (0)
#define HODE_DS_CP(CP) do { HODE_ND(HODE_DS(1, CP), HODE_DS(0, CP)); } while (0)
        HODE_DISPATCH_CP(F, a, st, HODE_DS_CP);
#undef HODE_DS_CP
#undef HODE_DS
        HODE_LAUNCH_CHECK();
        return 0;
    }
    if (!a_in.per_traj && a_in.pset == nullptr && a_in.batch >= 17 && a_in.batch <= 128 && pack_enabled()) {
        // mid-size groups: pack floor(256 / batch) of them into one CTA when that fills clearly more lanes than one group per CTA
        const SolveArgs& a = a_in;
        const int pthreads = CommPack::threads_for((int)a.batch);
        const int gpc = pthreads / (int)a.batch;
        const double util_pack = (double)(gpc * a.batch) / pthreads, util_cta = (double)a.batch / round_up32(a.batch);
        // measured (B200, 8 192 groups): batch 20 at D = 12 1.85 x faster packed (60 of 64 lanes instead of 20 of 32), batch 50 at
        // D = 6 10 % SLOWER packed (150 of 160 lanes instead of 50 of 64: the lock-step barrier over 5 warps and the wait for the
        // slowest of 3 groups cost more than the lanes gain) -> pack only for a clear gain in filled lanes
        if (gpc >= 2 && a.n_groups >= gpc && util_pack >= 1.4 * util_cta) {
            const int64_t nblk = (a.n_groups + gpc - 1) / gpc;
#define HODE_DP(ND, CP)                                                                                                        \
    do {                                                                                                                       \
        const size_t sh_ = ((CP ? 0 : round4(F::SP)) + CommPack::kFloats + D5Store<F>::fwd_floats(pthreads)) * sizeof(float); \
        int e_ = set_smem(dopri5_fwd_pack_kernel<F, ND, CP>, sh_);                                                             \
        if (e_ != 0) return e_;                                                                                                \
        dopri5_fwd_pack_kernel<F, ND, CP><<<(unsigned)nblk, pthreads, sh_, st>>>(a, gpc);                                      \
    } while (0)
#define HODE_DP_CP(CP) do { HODE_ND(HODE_DP(1, CP), HODE_DP(0, CP)); } while (0)
            HODE_DISPATCH_CP(F, a, st, HODE_DP_CP);
#undef HODE_DP_CP
#undef HODE_DP
            HODE_LAUNCH_CHECK();
            return 0;
        }
    }
    const SolveArgs a = a_in.per_traj ? flatten(a_in) : a_in;
    const int threads = a.per_traj ? (a.batch >= 128 ? 128 : round_up32(a.batch)) : round_up32(a.batch);
    const int tiles = a.per_traj ? (int)((a.batch + threads - 1) / threads) : 1;
    const int64_t nblk = a.n_groups * tiles;
#define HODE_DF(PT, ND, MAXT, CP)                                                                                             \
    do {                                                                                                                      \
        const size_t sh_ = ((CP ? 0 : round4(F::SP)) + 128 + D5Store<F>::fwd_floats(MAXT <= 128 ? 128 : threads)) * sizeof(float); \
        int e_ = set_smem(dopri5_fwd_kernel<F, PT, ND, MAXT, CP>, sh_);                                                       \
        if (e_ != 0) return e_;                                                                                               \
        dopri5_fwd_kernel<F, PT, ND, MAXT, CP><<<(unsigned)nblk, threads, sh_, st>>>(a, tiles);                               \
    } while (0)
#define HODE_DF_CP(CP)                                                                                       \
    do {                                                                                                     \
        if (a.per_traj) { HODE_ND(HODE_DF(true, 1, 128, CP), HODE_DF(true, 0, 128, CP)); }              \
        else if (threads <= 128) { HODE_ND(HODE_DF(false, 1, 128, CP), HODE_DF(false, 0, 128, CP)); }   \
        else { HODE_ND(HODE_DF(false, 1, HODE_DOPRI5_MAX_THREADS, CP), HODE_DF(false, 0, HODE_DOPRI5_MAX_THREADS, CP)); } \
    } while (0)
    HODE_DISPATCH_CP(F, a, st, HODE_DF_CP);
#undef HODE_DF_CP
#undef HODE_DF
    HODE_LAUNCH_CHECK();
    return 0;
}

template <class F>
int launch_dopri5_bwd(const hode_cfg& cfg, const SolveArgs& a_in, cudaStream_t st) {
    const SolveArgs a = (CoopOf<F>::value && !a_in.per_traj) ? a_in : flatten(a_in);
    const int threads = a.batch >= 128 ? 128 : round_up32(a.batch);
    const int tiles = (int)((a.batch + threads - 1) / threads);
    const int64_t nblk = a.n_groups * tiles;
    const bool nd1 = cfg.n_dose == 1;
    const bool eg = cfg.expert_grads != 0;
    size_t coop_floats = 0;
    if constexpr (CoopD5<F>::value) coop_floats = (size_t)CoopD5<F>::type::kStageFloats * 4;
#define HODE_DB(EG, ND, CP)                                                                                              \
    do {                                                                                                                 \
        const size_t sh_ = ((CP ? 0 : round4(F::SP)) + round4(F::P) + D5Store<F>::bwd_floats(128) + coop_floats) * sizeof(float); \
        int e_ = set_smem(dopri5_bwd_kernel<F, EG, ND, CP>, sh_);                                                        \
        if (e_ != 0) return e_;                                                                                          \
        dopri5_bwd_kernel<F, EG, ND, CP><<<(unsigned)nblk, threads, sh_, st>>>(a, tiles);                                \
    } while (0)
#define HODE_DB_CP(CP)                                                         \
    do {                                                                       \
        if (eg) { HODE_ND(HODE_DB(true, 1, CP), HODE_DB(true, 0, CP)); }  \
        else    { HODE_ND(HODE_DB(false, 1, CP), HODE_DB(false, 0, CP)); } \
    } while (0)
    HODE_DISPATCH_CP(F, a, st, HODE_DB_CP);
#undef HODE_DB_CP
#undef HODE_DB
    HODE_LAUNCH_CHECK();
    return 0;
}

template <class F>
int launch_dopri5_adj(const hode_cfg& cfg, const SolveArgs& a_in, cudaStream_t st) {
    const bool nd1 = cfg.n_dose == 1;
    const bool eg = cfg.expert_grads != 0;
    if (!a_in.per_traj && a_in.batch > HODE_DOPRI5_MAX_THREADS) return -2;
    if (a_in.per_traj && CoopOf<F>::value) return -3;  // NeuralODE: cooperative accumulators need converged warps
    if (a_in.adj_mixed && (a_in.per_traj || !MixedOk<F>::value)) return -4;  // mixed norm: batch-coupled RocheODE, D <= 8
    const SolveArgs a = a_in.per_traj ? flatten(a_in) : a_in;
    const int threads = a.per_traj ? (a.batch >= 128 ? 128 : round_up32(a.batch)) : round_up32(a.batch);
    const int tiles = a.per_traj ? (int)((a.batch + threads - 1) / threads) : 1;
    const int64_t nblk = a.n_groups * tiles;
    size_t coop_floats = 0;
    if constexpr (CoopD5<F>::value) coop_floats = (size_t)CoopD5<F>::type::kStageFloats * (size_t)(threads / 32);
    if constexpr (MixedOk<F>::value) { if (a.adj_mixed) coop_floats = mixed_floats<F>(threads); }
    const size_t sh_base = (size_t)128 + round4(F::P) + (size_t)14 * F::D * threads + coop_floats;
#define HODE_DA(PT, EG, ND, MAXT, CP)                                                                              \
    do {                                                                                                           \
        const size_t sh_ = ((CP ? 0 : round4(F::SP)) + sh_base) * sizeof(float);                                   \
        if (sh_ > 227 * 1024) return -2;                                                                           \
        int e_ = set_smem(dopri5_adj_kernel<F, PT, EG, ND, MAXT, CP>, sh_);                                        \
        if (e_ != 0) return e_;                                                                                    \
        dopri5_adj_kernel<F, PT, EG, ND, MAXT, CP><<<(unsigned)nblk, threads, sh_, st>>>(a, tiles);                \
    } while (0)
#define HODE_DA_ND(PT, EG, MAXT, CP) do { HODE_ND(HODE_DA(PT, EG, 1, MAXT, CP), HODE_DA(PT, EG, 0, MAXT, CP)); } while (0)
#define HODE_DA_EG(PT, MAXT, CP) do { if (eg) HODE_DA_ND(PT, true, MAXT, CP); else HODE_DA_ND(PT, false, MAXT, CP); } while (0)
#define HODE_DA_CP(CP)                                                              \
    do {                                                                            \
        if (a.per_traj) HODE_DA_EG(true, 128, CP);                                  \
        else if (threads <= 128) HODE_DA_EG(false, 128, CP);                        \
        else HODE_DA_EG(false, HODE_DOPRI5_MAX_THREADS, CP);                        \
    } while (0)
    HODE_DISPATCH_CP(F, a, st, HODE_DA_CP);
#undef HODE_DA_CP
#undef HODE_DA_EG
#undef HODE_DA_ND
#undef HODE_DA
    HODE_LAUNCH_CHECK();
    return 0;
}

}  // namespace hode
