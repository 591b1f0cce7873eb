// hode_eval.cu -- the Monte-Carlo evaluation step that follows the batched solve:
//   training_utils.py:144-151  50 decoder solves per test chunk  -> ONE solve with n_groups = mc (solver kernels)
//   training_utils.py:157-177  properscoring.crps_ensemble in Python double / triple loops -> the kernels below
// CRPS of an equally weighted ensemble x_1..x_M against an observation y (what properscoring.crps_ensemble returns):
//   crps = mean_s |x_s - y|  -  1/(2 M^2) sum_{s,s'} |x_s - x_s'|
#include <cuda_runtime.h>
#include <stdint.h>

namespace hode {

constexpr int kMaxMc = 128;  // ensemble members kept in registers / shared memory per element

// sum_{s<s'} |x_s - x_s'| and sum_s |x_s - y| for M values held in shared memory by one thread each
__device__ __forceinline__ float crps_from_smem(const float* __restrict__ xs, int stride, int M, float y) {
    float s1 = 0.0f, s2 = 0.0f;
    for (int i = 0; i < M; ++i) {
        const float xi = xs[i * stride];
        s1 += fabsf(xi - y);
        float acc = 0.0f;
        for (int j = i + 1; j < M; ++j) acc += fabsf(xi - xs[j * stride]);
        s2 += acc;
    }
    const float inv = 1.0f / (float)M;
    return s1 * inv - s2 * inv * inv;  // the pair sum counts each unordered pair once: 2 * s2 / (2 M^2)
}

// generic: truth [n], forecasts at fc + i * stride_n + s * stride_mc
__global__ void __launch_bounds__(128) crps_ensemble_kernel(const float* __restrict__ truth, const float* __restrict__ fc,
                                                            int64_t n, int32_t M, int64_t stride_n, int64_t stride_mc,
                                                            float* __restrict__ out) {
    extern __shared__ float sx[];  // [M][128]: member-major, one column per thread (conflict-free)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int s = 0; s < M; ++s) sx[s * 128 + threadIdx.x] = fc[i * stride_n + s * stride_mc];
    out[i] = crps_from_smem(sx + threadIdx.x, 128, M, truth[i]);
}

// fused read-out + CRPS: predictions x_s = W[o] . h[t, s * batch + b] + bias[o] are never written to memory.
// One CTA per (t, b): the M latent vectors go to shared memory once, thread o owns observation o.
template <int D>
__global__ void __launch_bounds__(128) decode_crps_kernel(int32_t obs, int32_t n_t, int64_t batch, int32_t M,
                                                          const float* __restrict__ h, const float* __restrict__ W,
                                                          const float* __restrict__ bias, const float* __restrict__ x,
                                                          int64_t st, int64_t sb, int64_t so, float* __restrict__ crps) {
    extern __shared__ float sm[];
    float* sh = sm;                 // [M][D]
    float* sx = sm + M * D;         // [M][blockDim.x]
    const int64_t tb = blockIdx.x;  // t * batch + b
    const int64_t t = tb / batch, b = tb % batch;
    const int64_t n_traj = batch * M;
    for (int e = threadIdx.x; e < M * D; e += blockDim.x) {
        const int s = e / D, d = e % D;
        sh[e] = h[((int64_t)t * n_traj + (int64_t)s * batch + b) * D + d];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < obs; o += blockDim.x) {
        float w[D];
#pragma unroll
        for (int d = 0; d < D; ++d) w[d] = W[o * D + d];
        const float bo = bias[o];
        for (int s = 0; s < M; ++s) {
            float v = bo;
#pragma unroll
            for (int d = 0; d < D; ++d) v = fmaf(w[d], sh[s * D + d], v);
            sx[s * blockDim.x + threadIdx.x] = v;
        }
        const float y = x[t * st + b * sb + (int64_t)o * so];
        crps[tb * obs + o] = crps_from_smem(sx + threadIdx.x, blockDim.x, M, y);
    }
}

int launch_crps_ensemble(const float* truth, const float* fc, int64_t n, int32_t M, int64_t stride_n, int64_t stride_mc,
                         float* out, cudaStream_t st) {
    if (M > kMaxMc) return -1;
    const size_t sh = sizeof(float) * (size_t)M * 128;
    cudaError_t e = cudaFuncSetAttribute(crps_ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return (int)e;
    const int64_t blocks = (n + 127) / 128;
    crps_ensemble_kernel<<<(unsigned)blocks, 128, sh, st>>>(truth, fc, n, M, stride_n, stride_mc, out);
    return (int)cudaGetLastError();
}

template <int D>
static int launch_decode_crps_d(int32_t obs, int32_t n_t, int64_t batch, int32_t M, const float* h, const float* W,
                                const float* b, const float* x, int64_t st, int64_t sb, int64_t so, float* crps,
                                cudaStream_t stream) {
    const int threads = obs >= 128 ? 128 : ((obs + 31) / 32) * 32;
    const size_t sh = sizeof(float) * ((size_t)M * D + (size_t)M * threads);
    cudaError_t e = cudaFuncSetAttribute(decode_crps_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return (int)e;
    decode_crps_kernel<D><<<(unsigned)(n_t * batch), threads, sh, stream>>>(obs, n_t, batch, M, h, W, b, x, st, sb, so, crps);
    return (int)cudaGetLastError();
}

int launch_decode_crps(int32_t D, int32_t obs, int32_t n_t, int64_t batch, int32_t M, const float* h, const float* W,
                       const float* b, const float* x, int64_t st, int64_t sb, int64_t so, float* crps,
                       cudaStream_t stream) {
    if (M > kMaxMc) return -1;
    if ((int64_t)n_t * batch > 0x7fffffffLL) return -1;
    switch (D) {
        case 4: return launch_decode_crps_d<4>(obs, n_t, batch, M, h, W, b, x, st, sb, so, crps, stream);
        case 6: return launch_decode_crps_d<6>(obs, n_t, batch, M, h, W, b, x, st, sb, so, crps, stream);
        case 8: return launch_decode_crps_d<8>(obs, n_t, batch, M, h, W, b, x, st, sb, so, crps, stream);
        case 12: return launch_decode_crps_d<12>(obs, n_t, batch, M, h, W, b, x, st, sb, so, crps, stream);
        default: return -1;
    }
}

}  // namespace hode
