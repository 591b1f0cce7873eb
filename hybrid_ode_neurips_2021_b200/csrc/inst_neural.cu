// inst_neural.cu -- explicit instantiation of the NeuralODE kernels for one latent dimension (-DHODE_INST_D=<D>).
#include "hode_launch.cuh"

#ifndef HODE_INST_D
#error "compile with -DHODE_INST_D=<latent_dim>"
#endif

namespace hode {
template int launch_fixed_fwd<Neural<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_fixed_fwd_sse<Neural<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_fixed_bwd<Neural<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_fixed_adj<Neural<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_fwd<Neural<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_bwd<Neural<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_adj<Neural<HODE_INST_D>>(const hode_cfg&, const SolveArgs&, cudaStream_t);
}  // namespace hode
