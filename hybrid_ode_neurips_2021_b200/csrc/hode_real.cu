// hode_real.cu -- sm_100a kernels for the real-data vector fields (hode_real.cuh): dose tables, fixed-grid forward and
// reverse sweep (DecoderReal.forward, model.py:833-862, only ever runs euler / midpoint / rk4 on them: real.sh:9-17).
// One thread = one patient; the packed parameters (1 160 - 2 686 floats at the ICU sizes) are staged into shared
// memory; per-thread parameter-gradient accumulators live in local memory (L1-resident at the reference's batch of
// 100) and are reduced warp shuffle -> shared atomics -> one global atomic per parameter per CTA.
#include <cuda_runtime.h>

#include "hode_bodies.cuh"
#include "hode_real.cuh"

namespace hode {

constexpr int kRealMaxHidden = 64;

// ---- dose tables ---------------------------------------------------------------------------------------------------
// kind 0 (RocheODEReal): tab[0][n][b] = S[n], tab[1][n][b] = S1[n], n = 0..T   (kel = params[1])
// kind 1 (NeuralODEReal / 2nd): tab[0][n][b] = cumsum(a)[n], n = 0..T-1; row T = 0
__global__ void __launch_bounds__(128) real_dose_table_kernel(int kind, const float* __restrict__ action, int64_t stride_t,
                                                              int64_t stride_b, int32_t T, int64_t n_traj,
                                                              const float* __restrict__ params, float* __restrict__ tab) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_traj) return;
    real_dose_table_column(kind, action + b * stride_b, stride_t, T, n_traj, kind == 0 ? params[1] : 0.0f, tab, b);
}

struct RealArgs {
    SolveArgs a;
    const float* tab;  // dose tables
    int32_t T;         // table rows - 1
    int32_t hidden;
    int32_t P;
};

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
static int launch_real(bool bwd, int method, const RealArgs& r, cudaStream_t st) {
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

int real_param_count(int field, int Z, int H) {
    if (H < 1 || H > kRealMaxHidden) return -1;
    switch (field) {
        case HODE_FIELD_ROCHE_REAL:
            if (Z == 4) return RocheReal<4>::p_count(H);
            if (Z == 20) return RocheReal<20>::p_count(H);
            return -1;
        case HODE_FIELD_NEURAL_REAL:
            if (Z == 4) return NeuralReal<4, false>::p_count(H);
            if (Z == 20) return NeuralReal<20, false>::p_count(H);
            return -1;
        case HODE_FIELD_NEURAL_REAL_2ND:
            if (Z == 8) return NeuralReal<8, true>::p_count(H);
            if (Z == 40) return NeuralReal<40, true>::p_count(H);
            return -1;
        default: return -1;
    }
}

int launch_real_dose_table(int field, const float* action, int64_t stride_t, int64_t stride_b, int32_t T, int64_t n_traj,
                           const float* params, float* tab, cudaStream_t st) {
    const int kind = field == HODE_FIELD_ROCHE_REAL ? 0 : 1;
    const int64_t blocks = (n_traj + 127) / 128;
    real_dose_table_kernel<<<(unsigned)blocks, 128, 0, st>>>(kind, action, stride_t, stride_b, T, n_traj, params, tab);
    return (int)cudaGetLastError();
}

int launch_real_fixed(bool bwd, int field, int Z, int method, const RealArgs& r, cudaStream_t st) {
    switch (field) {
        case HODE_FIELD_ROCHE_REAL:
            if (Z == 4) return launch_real<RocheReal<4>, true>(bwd, method, r, st);
            if (Z == 20) return launch_real<RocheReal<20>, true>(bwd, method, r, st);
            return -1;
        case HODE_FIELD_NEURAL_REAL:
            if (Z == 4) return launch_real<NeuralReal<4, false>, false>(bwd, method, r, st);
            if (Z == 20) return launch_real<NeuralReal<20, false>, false>(bwd, method, r, st);
            return -1;
        case HODE_FIELD_NEURAL_REAL_2ND:
            if (Z == 8) return launch_real<NeuralReal<8, true>, false>(bwd, method, r, st);
            if (Z == 40) return launch_real<NeuralReal<40, true>, false>(bwd, method, r, st);
            return -1;
        default: return -1;
    }
}

}  // namespace hode
