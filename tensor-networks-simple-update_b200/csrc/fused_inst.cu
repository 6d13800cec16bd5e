// fused_inst.cu — instantiation of the on-chip persistent run kernel (run_fused.cuh).
#define RUN_FUSED_IMPL
#include "run_fused.cuh"

cudaError_t launch_run_fused_dispatch(const FusedLaunch &p, int svd_rows, int max_m, int max_n, int smem_bytes, cudaStream_t st,
                                      bool dry_run, int *blocks_per_sm) {
  return launch_run_fused(p, svd_rows, max_m, max_n, smem_bytes, st, dry_run, blocks_per_sm);
}
