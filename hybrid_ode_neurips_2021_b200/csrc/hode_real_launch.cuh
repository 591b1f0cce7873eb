// hode_real_launch.cuh -- kernels and launcher template of the real-data fields (see hode_real.cu); instantiated per
// (field, latent width) in inst_real.cu so that the combinations compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "hode_real.cuh"
#include "hode_real_args.cuh"

namespace hode {

template <class F>
__device__ __forceinline__ DoseTab make_tab(const RealArgs& r, int64_t idx, bool two) {
    DoseTab d;
    const int64_t n_traj = r.a.n_groups * r.a.batch;
    d.s = r.tab + idx;
    d.s1 = two ? r.tab + (int64_t)(r.T + 1) * n_traj + idx : nullptr;
    d.stride = n_traj;
    d.T = r.T;
    return d;
}

template <class F, int METHOD, bool TWO>
__global__ void __launch_bounds__(128) real_fixed_fwd_kernel(const RealArgs r) {
    extern __shared__ float smem[];
    for (int i = threadIdx.x; i < r.P; i += blockDim.x) smem[i] = r.a.params[i];
    __syncthreads();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= r.a.n_groups * r.a.batch) return;
    typename F::Params sp{smem, r.hidden};
    fixed_fwd_traj<F, METHOD>(r.a, sp, make_tab<F>(r, idx, TWO), idx);
}

template <class F, int METHOD, bool TWO>
__global__ void __launch_bounds__(128) real_fixed_bwd_kernel(const RealArgs r) {
    extern __shared__ float smem[];
    float* sred = smem + r.P;
    for (int i = threadIdx.x; i < r.P; i += blockDim.x) { smem[i] = r.a.params[i]; sred[i] = 0.0f; }
    __syncthreads();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float acc[F::p_count(kRealMaxHidden)];  // local memory (interleaved across the warp by the hardware)
#pragma unroll 1
    for (int i = 0; i < r.P; ++i) acc[i] = 0.0f;
    if (idx < r.a.n_groups * r.a.batch) {
        typename F::Params sp{smem, r.hidden};
        fixed_bwd_traj<F, METHOD, true>(r.a, sp, make_tab<F>(r, idx, TWO), idx, acc);
    }
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int i = 0; i < r.P; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(&sred[i], v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < r.P; i += blockDim.x) atomicAdd(&r.a.grad_params[i], sred[i]);
}

template <class F, bool TWO>
int launch_real(bool bwd, int method, const RealArgs& r, cudaStream_t st) {
    const int64_t n_traj = r.a.n_groups * r.a.batch;
    const int threads = n_traj >= 128 ? 128 : (int)(((n_traj + 31) / 32) * 32);
    const int64_t nblk = (n_traj + threads - 1) / threads;
    const size_t sh = sizeof(float) * (size_t)r.P * (bwd ? 2 : 1);
#define HODE_RL(M)                                                                                                     \
    do {                                                                                                               \
        if (bwd) {                                                                                                     \
            cudaFuncSetAttribute(real_fixed_bwd_kernel<F, M, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); \
            real_fixed_bwd_kernel<F, M, TWO><<<(unsigned)nblk, threads, sh, st>>>(r);                           \
        } else {                                                                                                       \
            cudaFuncSetAttribute(real_fixed_fwd_kernel<F, M, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); \
            real_fixed_fwd_kernel<F, M, TWO><<<(unsigned)nblk, threads, sh, st>>>(r);                                   \
        }                                                                                                              \
    } while (0)
    switch (method) {
        case HODE_EULER: HODE_RL(M_EULER); break;
        case HODE_MIDPOINT: HODE_RL(M_MIDPOINT); break;
        case HODE_RK4_38: HODE_RL(M_RK4_38); break;
        default: return -1;
    }
#undef HODE_RL
    return (int)cudaGetLastError();
}

}  // namespace hode
