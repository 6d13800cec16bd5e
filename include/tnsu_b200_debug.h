/* tnsu_b200_debug.h — profiling / debugging entry points of libtnsu_b200.so.  NOT part of the drop-in boundary
 * (include/tnsu_b200.h): nothing in the reference corresponds to them.  bench.py uses tnsu_profile_sweep for its
 * per-kernel-class breakdown, tests/ and tools/ use tnsu_debug_edge to compare intermediates with
 * oracle/device_model.py. */
#ifndef TNSU_B200_DEBUG_H
#define TNSU_B200_DEBUG_H

#include "tnsu_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One sweep with CUDA events between the kernel classes (synchronises after every launch: for profiling
 * only).  ms_out[3] = total device time of {K1 absorb+LQ, K2 theta+SVD, K3 recompose+store} launches. */
int tnsu_profile_sweep(tnsu_engine *e, int slot, double *ms_out);

/* Debug: run ONE edge update of point `point` and return the kernel's intermediates
 * (L_i, L_j as [M][K] row-major, theta singular values) for step-by-step comparison with
 * oracle/device_model.py.  Buffers are complex128 / float64 with capacity 64*64 / 64. */
int tnsu_debug_edge(tnsu_engine *e, int slot, int edge, int point, void *l_i_c128, void *l_j_c128,
                    double *sigma,
                    int32_t *info /* [16]: M_i K_i M_j K_j n_i n_j D_new svd_sweeps, then K2 cycle counts theta/jacobi/extract */);

#ifdef __cplusplus
}
#endif
#endif /* TNSU_B200_DEBUG_H */
