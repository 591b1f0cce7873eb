// hode_launch.cuh -- sm_100a kernels wrapping the per-trajectory bodies, and their launchers.
//
// Mapping: one thread = one trajectory, whole solve in one launch, state/stages in registers; the parameter set of the
// CTA's group is staged once into shared memory (broadcast LDS on every use); observations of the solution are
// written with 16/8-byte vector stores, D*4 contiguous bytes per lane (full 32 B sectors).  A CTA never spans two
// groups, so a CTA has exactly one parameter set and (dopri5, batch-coupled) one controller.
#pragma once
#include <cuda_runtime.h>

#include "hode_bodies.cuh"

#ifndef HODE_BWD_MINBLOCKS
#define HODE_BWD_MINBLOCKS 1
#endif
#ifndef HODE_DOPRI5_MAX_THREADS
#define HODE_DOPRI5_MAX_THREADS 512
#endif

namespace hode {

// ---- group sum across the CTA (batch-coupled controller) or nothing (per-trajectory) -------------------------
struct CommNone {
    __device__ __forceinline__ void sum2(float&, float&) {}
};
struct CommCta {
    float* red;  // 2 * 32 floats of shared memory
    int nwarps;
    __device__ __forceinline__ void sum2(float& a, float& b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (nwarps > 1) {
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            if (lane == 0) { red[2 * wid] = a; red[2 * wid + 1] = b; }
            __syncthreads();
            float sa = 0.0f, sb = 0.0f;
            for (int i = 0; i < nwarps; ++i) { sa += red[2 * i]; sb += red[2 * i + 1]; }
            __syncthreads();
            a = sa; b = sb;
        }
    }
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
template <class F, int METHOD, int ND>
__global__ void __launch_bounds__(128) fixed_fwd_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ float smem[];
    const Tile tl = tile_of(a, tiles_per_group);
    stage_params<F>(a, tl.group, smem);
    if (tl.b >= a.batch) return;
    const int64_t idx = tl.group * a.batch + tl.b;
    if (ND > 0) {
        const DoseReg<(ND > 0 ? ND : 1)> ds = load_dose_reg<(ND > 0 ? ND : 1)>(a, idx);
        fixed_fwd_traj<F, METHOD>(a, smem, ds, idx);
    } else {
        const DoseMem ds = load_dose_mem(a, idx);
        fixed_fwd_traj<F, METHOD>(a, smem, ds, idx);
    }
}

template <class F, int METHOD, bool EG, int ND>
__global__ void __launch_bounds__(128, HODE_BWD_MINBLOCKS) fixed_bwd_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ float smem[];
    float* sp = smem;
    float* sred = smem + F::SP;
    const Tile tl = tile_of(a, tiles_per_group);
    stage_params<F>(a, tl.group, sp);
    float acc[F::P];
    zero_acc<F>(acc);
    if (tl.b < a.batch) {
        const int64_t idx = tl.group * a.batch + tl.b;
        if (ND > 0) {
            const DoseReg<(ND > 0 ? ND : 1)> ds = load_dose_reg<(ND > 0 ? ND : 1)>(a, idx);
            fixed_bwd_traj<F, METHOD, EG>(a, sp, ds, idx, acc);
        } else {
            const DoseMem ds = load_dose_mem(a, idx);
            fixed_bwd_traj<F, METHOD, EG>(a, sp, ds, idx, acc);
        }
    }
    reduce_param_grads<F>(a, tl.group, acc, sred);
}

// MAXT: launch bound.  128 for per-trajectory control and small groups (up to 255 registers per thread);
// HODE_DOPRI5_MAX_THREADS for a batch-coupled group that needs a whole large CTA (128 registers per thread).
template <class F, bool PER_TRAJ, int ND, int MAXT>
__global__ void __launch_bounds__(MAXT) dopri5_fwd_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ float smem[];
    float* sp = smem;
    float* red = smem + F::SP;
    const Tile tl = tile_of(a, tiles_per_group);
    stage_params<F>(a, tl.group, sp);
    const bool valid = tl.b < a.batch;
    if (PER_TRAJ && !valid) return;
    const int64_t b = valid ? tl.b : (a.batch - 1);
    const int64_t idx = tl.group * a.batch + b;
    const int64_t ctrl = PER_TRAJ ? idx : tl.group;
    const bool leader = PER_TRAJ ? true : (threadIdx.x == 0);
    const float count = PER_TRAJ ? (float)F::D : (float)(a.batch * F::D);
    if (ND > 0) {
        const DoseReg<(ND > 0 ? ND : 1)> ds = load_dose_reg<(ND > 0 ? ND : 1)>(a, idx);
        if (PER_TRAJ) { CommNone cm; dopri5_fwd_traj<F>(a, cm, sp, ds, idx, valid, ctrl, leader, count); }
        else { CommCta cm{red, (int)(blockDim.x >> 5)}; dopri5_fwd_traj<F>(a, cm, sp, ds, idx, valid, ctrl, leader, count); }
    } else {
        const DoseMem ds = load_dose_mem(a, idx);
        if (PER_TRAJ) { CommNone cm; dopri5_fwd_traj<F>(a, cm, sp, ds, idx, valid, ctrl, leader, count); }
        else { CommCta cm{red, (int)(blockDim.x >> 5)}; dopri5_fwd_traj<F>(a, cm, sp, ds, idx, valid, ctrl, leader, count); }
    }
}

template <class F, bool EG, int ND>
__global__ void __launch_bounds__(128) dopri5_bwd_kernel(const SolveArgs a, int tiles_per_group) {
    extern __shared__ float smem[];
    float* sp = smem;
    float* sred = smem + F::SP;
    const Tile tl = tile_of(a, tiles_per_group);
    stage_params<F>(a, tl.group, sp);
    float acc[F::P];
    zero_acc<F>(acc);
    if (tl.b < a.batch) {
        const int64_t idx = tl.group * a.batch + tl.b;
        const int64_t ctrl = a.per_traj ? idx : tl.group;
        if (ND > 0) {
            const DoseReg<(ND > 0 ? ND : 1)> ds = load_dose_reg<(ND > 0 ? ND : 1)>(a, idx);
            dopri5_bwd_traj<F, EG>(a, sp, ds, idx, ctrl, acc);
        } else {
            const DoseMem ds = load_dose_mem(a, idx);
            dopri5_bwd_traj<F, EG>(a, sp, ds, idx, ctrl, acc);
        }
    }
    reduce_param_grads<F>(a, tl.group, acc, sred);
}

// ---------------------------------------------------------------------------------------------------------------
// launchers (explicitly instantiated per field in inst_*.cu)
// ---------------------------------------------------------------------------------------------------------------
inline int round_up32(int64_t n) { return (int)(((n + 31) / 32) * 32); }

#define HODE_LAUNCH_CHECK()                                   \
    do {                                                      \
        cudaError_t e_ = cudaGetLastError();                  \
        if (e_ != cudaSuccess) return (int)e_;                \
    } while (0)

template <class F>
int launch_fixed_fwd(const hode_cfg& cfg, const SolveArgs& a, cudaStream_t st) {
    const int threads = a.batch >= 128 ? 128 : round_up32(a.batch);
    const int tiles = (int)((a.batch + threads - 1) / threads);
    const int64_t nblk = a.n_groups * tiles;
    const size_t sh = F::SP * sizeof(float);
#define HODE_FF(M, ND) fixed_fwd_kernel<F, M, ND><<<(unsigned)nblk, threads, sh, st>>>(a, tiles)
    const bool nd1 = cfg.n_dose == 1;
    switch (cfg.method) {
        case HODE_EULER: if (nd1) HODE_FF(M_EULER, 1); else HODE_FF(M_EULER, 0); break;
        case HODE_MIDPOINT: if (nd1) HODE_FF(M_MIDPOINT, 1); else HODE_FF(M_MIDPOINT, 0); break;
        case HODE_RK4_38: if (nd1) HODE_FF(M_RK4_38, 1); else HODE_FF(M_RK4_38, 0); break;
        default: return -1;
    }
#undef HODE_FF
    HODE_LAUNCH_CHECK();
    return 0;
}

template <class F>
int launch_fixed_bwd(const hode_cfg& cfg, const SolveArgs& a, cudaStream_t st) {
    const int threads = a.batch >= 128 ? 128 : round_up32(a.batch);
    const int tiles = (int)((a.batch + threads - 1) / threads);
    const int64_t nblk = a.n_groups * tiles;
    const size_t sh = (F::SP + F::P) * sizeof(float);
    const bool nd1 = cfg.n_dose == 1;
    const bool eg = cfg.expert_grads != 0;
#define HODE_FB(M)                                                                                          \
    do {                                                                                                    \
        if (eg) { if (nd1) fixed_bwd_kernel<F, M, true, 1><<<(unsigned)nblk, threads, sh, st>>>(a, tiles);  \
                  else fixed_bwd_kernel<F, M, true, 0><<<(unsigned)nblk, threads, sh, st>>>(a, tiles); }    \
        else    { if (nd1) fixed_bwd_kernel<F, M, false, 1><<<(unsigned)nblk, threads, sh, st>>>(a, tiles); \
                  else fixed_bwd_kernel<F, M, false, 0><<<(unsigned)nblk, threads, sh, st>>>(a, tiles); }   \
    } while (0)
    switch (cfg.method) {
        case HODE_EULER: HODE_FB(M_EULER); break;
        case HODE_MIDPOINT: HODE_FB(M_MIDPOINT); break;
        case HODE_RK4_38: HODE_FB(M_RK4_38); break;
        default: return -1;
    }
#undef HODE_FB
    HODE_LAUNCH_CHECK();
    return 0;
}

template <class F>
int launch_dopri5_fwd(const hode_cfg& cfg, const SolveArgs& a, cudaStream_t st) {
    const bool nd1 = cfg.n_dose == 1;
    const size_t sh = (F::SP + 64) * sizeof(float);
    if (a.per_traj) {
        const int threads = a.batch >= 128 ? 128 : round_up32(a.batch);
        const int tiles = (int)((a.batch + threads - 1) / threads);
        const int64_t nblk = a.n_groups * tiles;
        if (nd1) dopri5_fwd_kernel<F, true, 1, 128><<<(unsigned)nblk, threads, sh, st>>>(a, tiles);
        else dopri5_fwd_kernel<F, true, 0, 128><<<(unsigned)nblk, threads, sh, st>>>(a, tiles);
    } else {
        if (a.batch > HODE_DOPRI5_MAX_THREADS) return -2;
        const int threads = round_up32(a.batch);
        if (threads <= 128) {
            if (nd1) dopri5_fwd_kernel<F, false, 1, 128><<<(unsigned)a.n_groups, threads, sh, st>>>(a, 1);
            else dopri5_fwd_kernel<F, false, 0, 128><<<(unsigned)a.n_groups, threads, sh, st>>>(a, 1);
        } else {
            if (nd1) dopri5_fwd_kernel<F, false, 1, HODE_DOPRI5_MAX_THREADS><<<(unsigned)a.n_groups, threads, sh, st>>>(a, 1);
            else dopri5_fwd_kernel<F, false, 0, HODE_DOPRI5_MAX_THREADS><<<(unsigned)a.n_groups, threads, sh, st>>>(a, 1);
        }
    }
    HODE_LAUNCH_CHECK();
    return 0;
}

template <class F>
int launch_dopri5_bwd(const hode_cfg& cfg, const SolveArgs& a, cudaStream_t st) {
    const int threads = a.batch >= 128 ? 128 : round_up32(a.batch);
    const int tiles = (int)((a.batch + threads - 1) / threads);
    const int64_t nblk = a.n_groups * tiles;
    const size_t sh = (F::SP + F::P) * sizeof(float);
    const bool nd1 = cfg.n_dose == 1;
    if (cfg.expert_grads) {
        if (nd1) dopri5_bwd_kernel<F, true, 1><<<(unsigned)nblk, threads, sh, st>>>(a, tiles);
        else dopri5_bwd_kernel<F, true, 0><<<(unsigned)nblk, threads, sh, st>>>(a, tiles);
    } else {
        if (nd1) dopri5_bwd_kernel<F, false, 1><<<(unsigned)nblk, threads, sh, st>>>(a, tiles);
        else dopri5_bwd_kernel<F, false, 0><<<(unsigned)nblk, threads, sh, st>>>(a, tiles);
    }
    HODE_LAUNCH_CHECK();
    return 0;
}

}  // namespace hode
