// hode_real_args.cuh -- argument block and launcher declaration shared by hode_api.cu, hode_real.cu and inst_real.cu
#pragma once
#include <cuda_runtime.h>

#include "hode_bodies.cuh"

namespace hode {

constexpr int kRealMaxHidden = 64;

struct RealArgs {
    SolveArgs a;
    const float* tab;  // dose tables
    int32_t T;         // table rows - 1
    int32_t hidden;
    int32_t P;
};

template <class F, bool TWO>
int launch_real(bool bwd, int method, const RealArgs& r, cudaStream_t st);  // defined in hode_real_launch.cuh

}  // namespace hode
