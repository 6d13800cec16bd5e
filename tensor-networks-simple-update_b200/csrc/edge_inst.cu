// edge_inst.cu — explicit instantiations of the edge-update kernel variants.
// Compiled once per (scalar type, MMAX): nvcc -DINST_CPLX=0|1 -DINST_MM=8|16|32 (see __graft_entry__.build()).
#include "edge_update.cuh"

#if INST_CPLX
typedef cplx inst_scalar;
#else
typedef double inst_scalar;
#endif

#define INSTANTIATE(NC, CL)                                                                                      \
  template cudaError_t launch_edge_variant<inst_scalar, INST_MM, NC, CL>(const EdgeLaunch<inst_scalar> &, int, int, \
                                                                         int, cudaStream_t);
INSTANTIATE(1, false)
INSTANTIATE(1, true)
#if !(INST_CPLX && INST_MM == 32)  // complex128 32x2 columns would need 256 data registers per thread
INSTANTIATE(2, false)
INSTANTIATE(2, true)
#endif
