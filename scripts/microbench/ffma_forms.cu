// Microbenchmark (development aid): FP32 FMA issue forms on sm_100a -- 3-register FFMA, constant-operand FFMA, packed FFMA2
// (fma.rn.f32x2).  Decides whether packing the per-trajectory FMA work of the solver kernels into FFMA2 pays.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/ffma_forms scripts/microbench/ffma_forms.cu
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int NACC = 16;

template <int MODE>
__global__ void __launch_bounds__(256) probe(const float* __restrict__ in, float* out, int iters, float pa, float pb) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = in[(threadIdx.x + i) & 255];
    float ra = in[256 + (threadIdx.x & 1)], rb = in[258 + (threadIdx.x & 1)];  // runtime values -> registers
    if (MODE == 0) {  // 3-register form
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], ra, rb);
        }
    } else if (MODE == 1) {  // constant-bank operands (kernel parameters)
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], pa, pb);
        }
    } else if (MODE == 2) {  // packed: 8 FFMA2 per 16 accumulators
        float2 a2 = make_float2(ra, ra), b2 = make_float2(rb, rb);
        float2 v[NACC / 2];
#pragma unroll
        for (int i = 0; i < NACC / 2; ++i) v[i] = make_float2(acc[2 * i], acc[2 * i + 1]);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC / 2; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
        }
#pragma unroll
        for (int i = 0; i < NACC / 2; ++i) { acc[2 * i] = v[i].x; acc[2 * i + 1] = v[i].y; }
    } else if (MODE == 3) {  // accumulate form acc += u * y[i]  (3 distinct registers, the vjp's dW update)
        float y[NACC];
#pragma unroll
        for (int i = 0; i < NACC; ++i) y[i] = in[(threadIdx.x + 7 * i) & 255];
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(ra, y[i], acc[i]);
            ra += rb;
        }
    } else if (MODE == 4) {  // same, packed
        float2 y2[NACC / 2], v[NACC / 2];
#pragma unroll
        for (int i = 0; i < NACC / 2; ++i) {
            y2[i] = make_float2(in[(threadIdx.x + 14 * i) & 255], in[(threadIdx.x + 14 * i + 7) & 255]);
            v[i] = make_float2(acc[2 * i], acc[2 * i + 1]);
        }
        for (int it = 0; it < iters; ++it) {
            const float2 u2 = make_float2(ra, ra);
#pragma unroll
            for (int i = 0; i < NACC / 2; ++i) v[i] = __ffma2_rn(u2, y2[i], v[i]);
            ra += rb;
        }
#pragma unroll
        for (int i = 0; i < NACC / 2; ++i) { acc[2 * i] = v[i].x; acc[2 * i + 1] = v[i].y; }
    } else if (MODE == 5) {  // constant operand + packed: weights from the constant bank as a pair
        float2 v[NACC / 2];
#pragma unroll
        for (int i = 0; i < NACC / 2; ++i) v[i] = make_float2(acc[2 * i], acc[2 * i + 1]);
        const float2 a2 = make_float2(pa, pb), b2 = make_float2(pb, pa);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC / 2; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
        }
#pragma unroll
        for (int i = 0; i < NACC / 2; ++i) { acc[2 * i] = v[i].x; acc[2 * i + 1] = v[i].y; }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, const float* in, float* out, int sms) {
    const int iters = 1 << 14, blocks = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<MODE><<<blocks, 256>>>(in, out, iters, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        probe<MODE><<<blocks, 256>>>(in, out, iters, 0.999f, 0.001f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double fma = (double)blocks * 256 * NACC * (double)iters;
    printf("{\"mode\": \"%s\", \"ms\": %.4f, \"TFLOPs\": %.2f, \"fma_per_clk_per_sm_at_1965MHz\": %.1f}\n", name, best,
           2 * fma / (best * 1e-3) / 1e12, fma / (best * 1e-3) / 1.965e9 / sms);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float h[512]; for (int i = 0; i < 512; ++i) h[i] = 0.5f + 0.001f * i;
    float *in, *out; cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, 4); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("ffma_3reg", in, out, p.multiProcessorCount);
    run<1>("ffma_const_operands", in, out, p.multiProcessorCount);
    run<2>("ffma2_3reg", in, out, p.multiProcessorCount);
    run<3>("ffma_accumulate_u_y", in, out, p.multiProcessorCount);
    run<4>("ffma2_accumulate_u_y", in, out, p.multiProcessorCount);
    run<5>("ffma2_const_pair", in, out, p.multiProcessorCount);
    return 0;
}
