// lqc_inst.cu — instantiation of the streaming CholeskyQR2 K1 / K3 kernels (lq_chol.cuh).
#include "lq_chol.cuh"

cudaError_t launch_lq_chol_dispatch(const EdgeLaunch<double> &p, int n_items, int m_max, int *flags, cudaStream_t st) {
  return launch_lq_chol(p, n_items, m_max, flags, st);
}
cudaError_t launch_rc_chol_dispatch(const EdgeLaunch<double> &p, int n_items, int m_max, const int *flags, cudaStream_t st) {
  return launch_rc_chol(p, n_items, m_max, flags, st);
}
cudaError_t launch_lq_chol_cx_dispatch(const EdgeLaunch<cplx> &p, int n_items, int m_max, int *flags, cudaStream_t st) {
  return launch_lq_chol_cx(p, n_items, m_max, flags, st);
}
cudaError_t launch_rc_chol_cx_dispatch(const EdgeLaunch<cplx> &p, int n_items, int m_max, const int *flags, cudaStream_t st) {
  return launch_rc_chol_cx(p, n_items, m_max, flags, st);
}
