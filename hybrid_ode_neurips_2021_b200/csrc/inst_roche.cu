// inst_roche.cu -- explicit instantiation of the RocheODE kernels for one latent dimension and one Hill-exponent
// variant (-DHODE_INST_D=<D> -DHODE_INST_HILL2=<0|1>).  One translation unit per combination so that they compile in
// parallel.
#include "hode_launch.cuh"

#if !defined(HODE_INST_D) || !defined(HODE_INST_HILL2)
#error "compile with -DHODE_INST_D=<latent_dim> -DHODE_INST_HILL2=<0|1>"
#endif

namespace hode {
using InstField = Roche<HODE_INST_D, (HODE_INST_HILL2 != 0)>;
template int launch_fixed_fwd<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_fixed_bwd<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_fwd<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_bwd<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
}  // namespace hode
