// rc_inst.cu — instantiation of the tensor-core recompose kernel variants (recompose_dmma.cuh).
#include "recompose_dmma.cuh"

cudaError_t launch_recompose_dmma_dispatch(const EdgeLaunch<double> &p, int n_items, int mout_max, int k_max, int ncl_cap,
                                           cudaStream_t st) {
  return launch_recompose_dmma(p, n_items, mout_max, k_max, ncl_cap, st);
}
