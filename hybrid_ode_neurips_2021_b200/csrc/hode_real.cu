// hode_real.cu -- sm_100a kernels for the real-data vector fields (hode_real.cuh): dose tables, fixed-grid forward and
// reverse sweep (DecoderReal.forward, model.py:833-862, only ever runs euler / midpoint / rk4 on them: real.sh:9-17).
// One thread = one patient; the packed parameters (1 160 - 2 686 floats at the ICU sizes) are staged into shared
// memory; per-thread parameter-gradient accumulators live in local memory (L1-resident at the reference's batch of
// 100) and are reduced warp shuffle -> shared atomics -> one global atomic per parameter per CTA.
#include <cuda_runtime.h>

#include "hode_real.cuh"
#include "hode_real_args.cuh"

namespace hode {

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
