// inst_roche.cu -- explicit instantiation of the RocheODE kernels for one latent dimension (-DHODE_INST_D=<D>).
// One translation unit per D so that the four dimensions compile in parallel.
#include "hode_launch.cuh"

#ifndef HODE_INST_D
#error "compile with -DHODE_INST_D=<latent_dim>"
#endif

namespace hode {
template int launch_fixed_fwd<Roche<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_fixed_bwd<Roche<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_fwd<Roche<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_bwd<Roche<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
}  // namespace hode
