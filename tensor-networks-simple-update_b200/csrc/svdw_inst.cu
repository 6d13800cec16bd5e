// svdw_inst.cu — instantiation of the register-resident warp SVD kernel variants (svd_warp.cuh).
#include "svd_warp.cuh"

cudaError_t launch_svd_warp_dispatch(const EdgeLaunch<double> &p, int n_items, cudaStream_t st) {
  return launch_svd_warp(p, n_items, st);
}
