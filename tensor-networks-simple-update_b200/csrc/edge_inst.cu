// edge_inst.cu — explicit instantiations of the edge-update kernel variants.
// Compiled once per (scalar type, MMAX): nvcc -DINST_CPLX=0|1 -DINST_MM=8|16|32 (see tnsu_b200/_build.py).
#include "edge_update.cuh"

#if INST_CPLX
typedef cplx inst_scalar;
#else
typedef double inst_scalar;
#endif

#define INSTANTIATE(NC)                                                                                          \
  template cudaError_t launch_lq_variant<inst_scalar, INST_MM, NC>(const EdgeLaunch<inst_scalar> &, int, cudaStream_t); \
  template cudaError_t launch_rc_variant<inst_scalar, INST_MM, NC>(const EdgeLaunch<inst_scalar> &, int, cudaStream_t);
INSTANTIATE(1)
#if !(INST_CPLX && INST_MM == 32)  // complex128 32x2 columns would need 256 data registers per thread
INSTANTIATE(2)
#endif
#if INST_MM == 8  // the SVD kernel has no MMAX parameter: build it once per scalar type
template cudaError_t launch_svd<inst_scalar>(const EdgeLaunch<inst_scalar> &, int, cudaStream_t);
#endif
