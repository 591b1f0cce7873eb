// hode_sse.cuh -- output sinks that consume the latent solution at an output time instead of storing it:
// read-out x_hat = W h + b (model.py:1097-1100, 1120), masked squared error (model.py:1179) and ALL its gradients
//   loss   += sum_o (x - x_hat)^2 mask / n_norm
//   grad_h  = W^T c,  grad_W += c h^T,  grad_b += c      with c = -2 / n_norm (x - x_hat) mask
// (the loss is a scalar, so d loss / d h(t_j) depends on h(t_j) alone and is known the moment h(t_j) exists).
// Fused into the forward solve, the latent solution is never written or re-read, the separate decode pass over x / mask
// disappears from the training step, and the x / mask stream (384 B per trajectory and output time at the C2 shape) hides
// under an FMA-bound kernel.
#pragma once
#include "hode_bodies.cuh"

namespace hode {

// ---- host / emulation sink: plain loops (tests/hostsim) ------------------------------------------------------------------
template <int D>
struct SseSinkHost {
    const SolveArgs* a;
    int64_t n_traj, idx;
    float* gw;   // [obs * D + obs] accumulators of the caller (grad_W row-major, then grad_b)
    double loss;
    HODE_HD void emit(int j, const float (&v)[D]) {
        const int obs = a->sse_obs;
        const float* xr = a->sse_x + ((int64_t)j * n_traj + idx) * obs;
        const float* mr = a->sse_mask + ((int64_t)j * n_traj + idx) * obs;
        float gh[D];
        for (int d = 0; d < D; ++d) gh[d] = 0.0f;
        for (int o = 0; o < obs; ++o) {
            float xh = a->sse_b[o];
            for (int d = 0; d < D; ++d) xh = fmaf(a->sse_w[o * D + d], v[d], xh);
            const float diff = xr[o] - xh;
            const float dm = diff * mr[o];
            loss += (double)(diff * dm);
            const float c = a->sse_scale * dm;
            for (int d = 0; d < D; ++d) {
                gh[d] = fmaf(c, a->sse_w[o * D + d], gh[d]);
                gw[o * D + d] = fmaf(c, v[d], gw[o * D + d]);
            }
            gw[obs * D + o] += c;
        }
        if (a->sse_grad_h != nullptr) store_vec<D>(a->sse_grad_h + ((int64_t)j * n_traj + idx) * D, gh);
        if (a->h_out != nullptr) store_vec<D>(a->h_out + ((int64_t)j * n_traj + idx) * D, v);
    }
};

#if HODE_DEVICE_BUILD && defined(__CUDACC__)
// ---- device sink -------------------------------------------------------------------------------------------------------
// Read-out weights in the constant bank (one weight set per launch), two layouts like ml_net's:
//   WR [obs][D]        row-major: pairs over d for grad_h += c_o W[o, :]
//   WP [obs/2][D][2]   interleaved by observation pair: pairs over o for x_hat
//   B  [obs]
constexpr int kReadoutMaxObs = 128, kReadoutMaxD = 8;
constexpr int kReadoutFloats = 2 * kReadoutMaxObs * kReadoutMaxD + kReadoutMaxObs;
static __constant__ __align__(16) float c_readout[kReadoutFloats];
static __device__ __align__(16) float g_readout_stage[kReadoutFloats];

static __global__ void __launch_bounds__(128) prep_readout_kernel(const float* __restrict__ W, const float* __restrict__ b, int obs, int D) {
    for (int i = threadIdx.x; i < obs * D; i += blockDim.x) {
        const int o = i / D, d = i % D;
        const float w = W[i];
        g_readout_stage[i] = w;
        g_readout_stage[obs * D + ((o >> 1) * D + d) * 2 + (o & 1)] = w;
    }
    for (int o = threadIdx.x; o < obs; o += blockDim.x) g_readout_stage[2 * obs * D + o] = b[o];
}

__device__ __forceinline__ unsigned sse_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sse_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sse_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sse_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sse_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sse_mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(sse_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void sse_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sse_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(sse_smem_u32(bar)) : "memory");
}

// One sink per thread; the 32 lanes of a warp share a tile of 32 consecutive trajectories.  Per output time t_j the
// warp's x and mask rows (32 x OBS floats each, contiguous in global memory) arrive by ONE TMA bulk copy each, requested
// by lane 0 right after the previous output time was consumed -- a whole output interval (16 solver steps at the C2 shape)
// ahead of their use.  emit() must be called by all 32 lanes together (fixed grids: uniform trip counts).
//   main pass   lane = trajectory: x_hat, loss term, c (written over the x tile), grad_h = W^T c
//   owner pass  lane = (observation group og = lane / D, state dimension d = lane % D): grad_W[o, d] += sum_tt c[tt, o] h[tt, d]
//               for its OPL observations over the warp's 32 trajectories; lane = observation (+ 32 k): grad_b
// OBS is a compile-time constant (20 / 24 / 40 / 80: the reference's observation widths): every constant-bank and
// shared-memory offset is an immediate.  emit() is branch-free over the lanes (rows beyond the cohort are masked by
// selects), so the compiler may keep the weights on the uniform datapath like the solver loop does.
template <int D, int OBS>
struct SseSink {
    static constexpr int OG = 32 / D;                                 // observation groups of the owner pass
    static constexpr int OPL = ((OBS + OG - 1) / OG + 1) / 2 * 2;     // observations per owner lane (even)
    static constexpr int NQ = OBS / 4;
    static constexpr int kAccFloats = 1 + OPL + 4;                    // per thread: loss, grad_W block, grad_b entries
    static constexpr int kTileFloats = 2 * 32 * OBS;                  // per warp: x tile, mask tile
    static_assert(OBS % 4 == 0 && D % 2 == 0 && D <= kReadoutMaxD && OBS <= kReadoutMaxObs, "packed loops / constant array");
    // offsets (floats) into the CTA's dynamic shared memory: x tile (-> c), mask tile (its head doubles as the published
    // h rows [32][D] of the owner pass), this thread's accumulators [kAccFloats][128], the warp's mbarrier
    int off_x, off_acc, off_bar;
    unsigned parity;
    int rows;  // valid trajectories of this warp's tile
    int n_t;
    int64_t n_traj, idx0;
    const float* x;
    const float* mask;
    float* grad_h;
    float* h_out;
    float scale;
    bool want_w;

    __device__ __forceinline__ void init(const SolveArgs& args, int tile_off, int acc_off, int bar_off, int64_t first_traj,
                                         int64_t total) {
        extern __shared__ __align__(16) float smem[];
        x = args.sse_x; mask = args.sse_mask; grad_h = args.sse_grad_h; h_out = args.h_out; scale = args.sse_scale;
        n_t = args.n_t; want_w = args.sse_grad_w != nullptr; n_traj = total; idx0 = first_traj;
        off_x = tile_off; off_acc = acc_off + (int)threadIdx.x; off_bar = bar_off; parity = 0;
        const int64_t left = total - first_traj;
        rows = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
#pragma unroll
        for (int e = 0; e < kAccFloats; ++e) smem[off_acc + e * 128] = 0.0f;
        if ((threadIdx.x & 31) == 0) {
            sse_mbar_init(reinterpret_cast<uint64_t*>(smem + off_bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        request(0);
    }
    __device__ __forceinline__ void request(int j) {
        extern __shared__ __align__(16) float smem[];
        if ((threadIdx.x & 31) != 0 || rows == 0) return;
        const unsigned bytes = (unsigned)rows * (unsigned)OBS * 4u;
        const int64_t g = ((int64_t)j * n_traj + idx0) * OBS;
        uint64_t* bar = reinterpret_cast<uint64_t*>(smem + off_bar);
        sse_mbar_expect_tx(bar, 2u * bytes);
        sse_bulk_g2s(smem + off_x, x + g, bytes, bar);
        sse_bulk_g2s(smem + off_x + 32 * OBS, mask + g, bytes, bar);
    }
    // (fixed_fwd_traj has a single emit site, outside its stepping loop; the per-thread accumulators live in shared memory:
    // the solver loop keeps the register allocation and the uniform-register weight operands of the plain forward kernel)
    __device__ __forceinline__ void emit(int j, const float (&v)[D]) {
        extern __shared__ __align__(16) float smem[];
        const int lane = threadIdx.x & 31;
        const bool valid = lane < rows;
        float* sx = smem + off_x;
        float* sm = sx + 32 * OBS;
        float* acc = smem + off_acc;
        if (rows > 0) sse_mbar_wait(reinterpret_cast<uint64_t*>(smem + off_bar), parity);  // warp-uniform
        parity ^= 1u;
        float4* xr = reinterpret_cast<float4*>(sx + lane * OBS);
        const float4* mr = reinterpret_cast<const float4*>(sm + lane * OBS);
        const float sc = valid ? scale : 0.0f;  // rows beyond the cohort hold stale shared memory: they must add nothing
        float lsum = 0.0f;
        float gh[D];
#pragma unroll
        for (int d = 0; d < D; ++d) gh[d] = 0.0f;
        constexpr int WR = 0, WP = OBS * D, BB = 2 * OBS * D;
#pragma unroll 1
        for (int q = 0; q < NQ; ++q) {
            const float4 x4 = xr[q], m4 = mr[q];
            const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ms[4] = {m4.x, m4.y, m4.z, m4.w};
            float xh[4], cs[4];
#pragma unroll
            for (int p = 0; p < 2; ++p) {  // observation pairs (4q + 2p, 4q + 2p + 1)
                const int op = 2 * q + p;
                float a0 = c_readout[BB + 2 * op], a1 = c_readout[BB + 2 * op + 1];
#pragma unroll
                for (int d = 0; d < D; ++d)
                    fma2s(v[d], c_readout[WP + (op * D + d) * 2], c_readout[WP + (op * D + d) * 2 + 1], a0, a1, a0, a1);
                xh[2 * p] = a0; xh[2 * p + 1] = a1;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float diff = xs[e] - xh[e];
                const float dm = diff * ms[e];
                lsum = fmaf(diff, dm, lsum);
                cs[e] = valid ? sc * dm : 0.0f;  // (select, not a product: a stale row may hold NaN)
                const int o = 4 * q + e;
#pragma unroll
                for (int d = 0; d < D; d += 2)
                    fma2s(cs[e], c_readout[WR + o * D + d], c_readout[WR + o * D + d + 1], gh[d], gh[d + 1], gh[d], gh[d + 1]);
            }
            xr[q] = make_float4(cs[0], cs[1], cs[2], cs[3]);
        }
        if (valid) {
            acc[0] += lsum;
            const int64_t row = (int64_t)j * n_traj + idx0 + lane;
            if (grad_h != nullptr) store_vec<D>(grad_h + row * D, gh);
            if (h_out != nullptr) store_vec<D>(h_out + row * D, v);
        }
        __syncwarp();  // every lane is done with the mask tile: its head becomes the h rows
#pragma unroll
        for (int d = 0; d < D; ++d) sm[lane * D + d] = valid ? v[d] : 0.0f;
        __syncwarp();
        if (want_w) {  // warp-uniform
            const int og = lane / D, d = lane - og * D;
            if (OG * D == 32 || og < OG) {
                const float* cp = sx + og * OPL;
                const int o_lim = OBS - og * OPL;  // observations of this group that exist
                float gw[OPL];
#pragma unroll
                for (int e = 0; e < OPL; ++e) gw[e] = acc[(1 + e) * 128];
#pragma unroll 4
                for (int tt = 0; tt < 32; ++tt) {
                    const float hv = sm[tt * D + d];
#pragma unroll
                    for (int e = 0; e < OPL; e += 2) {
                        if (e < o_lim) {  // OBS and OPL are even: pairs never straddle the end
                            const float2 c2 = *reinterpret_cast<const float2*>(cp + tt * OBS + e);
                            fma2s(hv, c2.x, c2.y, gw[e], gw[e + 1], gw[e], gw[e + 1]);
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < OPL; ++e) acc[(1 + e) * 128] = gw[e];
            }
#pragma unroll
            for (int k = 0; k < (OBS + 31) / 32; ++k) {
                const int o = lane + 32 * k;
                if (o < OBS) {
                    float s = 0.0f;
#pragma unroll 8
                    for (int tt = 0; tt < 32; ++tt) s += sx[tt * OBS + o];
                    acc[(1 + OPL + k) * 128] += s;
                }
            }
        }
        __syncwarp();  // tiles consumed: request the next output time's rows
        if (j + 1 < n_t) {
            if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            request(j + 1);
        }
    }
    // owned partial sums -> shared accumulator `sred` [OBS * D + OBS] (zeroed by the caller) ; loss -> global
    __device__ __forceinline__ void flush(float* sred, float* loss, float inv_norm) {
        extern __shared__ __align__(16) float smem[];
        const int lane = threadIdx.x & 31;
        const float* acc = smem + off_acc;
        if (want_w) {
            const int og = lane / D, d = lane - og * D;
            if (og < OG) {
#pragma unroll
                for (int e = 0; e < OPL; ++e) {
                    const int o = og * OPL + e;
                    if (o < OBS) atomicAdd(&sred[o * D + d], acc[(1 + e) * 128]);
                }
            }
#pragma unroll
            for (int k = 0; k < (OBS + 31) / 32; ++k) {
                const int o = lane + 32 * k;
                if (o < OBS) atomicAdd(&sred[OBS * D + o], acc[(1 + OPL + k) * 128]);
            }
        }
        float l = acc[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        if (lane == 0) atomicAdd(loss, l * inv_norm);
    }
};
#endif  // device

}  // namespace hode
