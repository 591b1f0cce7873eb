// inst_roche.cu -- explicit instantiation of the RocheODE kernels for one latent dimension and one Hill-exponent
// variant (-DHODE_INST_D=<D> -DHODE_INST_HILL2=<0|1>).  One translation unit per combination so that they compile in
// parallel.
#include "hode_launch.cuh"

#if !defined(HODE_INST_D) || !defined(HODE_INST_HILL2)
#error "compile with -DHODE_INST_D=<latent_dim> -DHODE_INST_HILL2=<0|1|2>  (2 = the ablation field, model.py:545-549)"
#endif

namespace hode {
#if HODE_INST_HILL2 == 2
using InstField = Roche<HODE_INST_D, true, true>;
#else
using InstField = Roche<HODE_INST_D, (HODE_INST_HILL2 != 0)>;
#endif
template int launch_fixed_fwd<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_fixed_fwd_sse<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_fixed_bwd<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_fixed_adj<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_fwd<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_bwd<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
template int launch_dopri5_adj<InstField>(const hode_cfg&, const SolveArgs&, cudaStream_t);
}  // namespace hode
