// inst_real.cu -- explicit instantiation of the real-data kernels for one (field, latent width):
//   -DHODE_REAL_FIELD=<2 RocheReal | 3 NeuralReal | 4 NeuralReal2nd> -DHODE_REAL_Z=<latent width>
#include "hode_real_launch.cuh"

#if !defined(HODE_REAL_FIELD) || !defined(HODE_REAL_Z)
#error "compile with -DHODE_REAL_FIELD=<2|3|4> -DHODE_REAL_Z=<latent_dim>"
#endif

namespace hode {
#if HODE_REAL_FIELD == 2
template int launch_real<RocheReal<HODE_REAL_Z>, true>(bool, int, const RealArgs&, cudaStream_t);
#elif HODE_REAL_FIELD == 3
template int launch_real<NeuralReal<HODE_REAL_Z, false>, false>(bool, int, const RealArgs&, cudaStream_t);
#else
template int launch_real<NeuralReal<HODE_REAL_Z, true>, false>(bool, int, const RealArgs&, cudaStream_t);
#endif
}  // namespace hode
