/* tnsu_b200.h — C ABI of the B200-native Simple-Update engine (libtnsu_b200.so).
 *
 * Drop-in boundary for the hot path of RoyElkabetz/Tensor-Networks-Simple-Update (tnsu 1.0.4).
 * The reference has no FFI: its boundary is the Python method surface of `SimpleUpdate`
 * (src/tnsu/simple_update.py:12-667) over the attribute surface of `TensorNetwork`
 * (src/tnsu/tensor_network.py:18-356).  Each entry point below names the reference code it
 * replaces; `INTEGRATION.md` shows the ctypes stub a tnsu maintainer would add.
 *
 * Conventions
 *  - every call returns 0 on success or a negative TNSU_E_* code; tnsu_last_error() gives the text.
 *    No exception or abort crosses this boundary.
 *  - tensors are C-contiguous, axis 0 physical (size d), then one axis per structure-matrix entry of
 *    that row in axis-number order (tensor_network.py:81-105); weights are float64 vectors; the
 *    structure matrix is int64 n x m row-major, entries = 1-based axis numbers, 0 = not connected
 *    (structure_matrix_constructor.py:148-178).
 *  - an engine holds `batch` independent networks ("points") of ONE structure and ONE shape
 *    trajectory; batched buffers are [point][...] with the point index outermost.
 *  - gates / Hamiltonians cross the boundary as complex128 (interleaved re,im) regardless of the
 *    engine dtype; an f64 engine keeps the real parts (the caller guarantees imag == 0).
 *  - the library copies inputs during the call and never keeps a caller pointer.
 *  - a handle is bound to one device and one stream; calls on one handle must be serialised by the
 *    caller; distinct handles are independent.  There is no CPU fallback: without a CUDA device
 *    tnsu_create fails with TNSU_E_CUDA.
 */
#ifndef TNSU_B200_H
#define TNSU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tnsu_engine tnsu_engine;

enum { TNSU_F64 = 0, TNSU_C128 = 1 };

enum {
  TNSU_OK = 0,
  TNSU_E_ARG = -1,         /* bad argument / shape (AssertionError / ValueError in the reference) */
  TNSU_E_UNSUPPORTED = -2, /* dims exceed what the sm_100a kernels hold on chip */
  TNSU_E_CUDA = -3,        /* CUDA runtime error, or no device */
  TNSU_E_NONFINITE = -4,   /* NaN/Inf detected in results */
  TNSU_E_STATE = -5        /* call order (state not uploaded, gates missing, ...) */
};

/* ABI version of this header; tnsu_abi_version() must return the same number. */
#define TNSU_ABI_VERSION 2
int tnsu_abi_version(void);

/* Text of the last error on this handle (or of the last failed tnsu_create when e == NULL). */
const char *tnsu_last_error(const tnsu_engine *e);

/* Host-only scheduler query (no GPU needed): dependency level of every edge for one sweep in the
 * reference's edge order (simple_update.py:219): an edge waits for every EARLIER edge that shares a
 * tensor with it, so executing level by level reproduces the sequential sweep exactly.
 * Returns the number of levels, or a negative error code for an invalid structure matrix. */
int tnsu_schedule_levels(int n, int m, const int64_t *smat, int32_t *level_out /* [m] */);

/* Create an engine for `batch` networks of structure `smat` (n tensors x m edges), physical
 * dimension d, initial bond dimensions edge_dims[m] and truncation d_max.
 * Replaces: TensorNetwork.__init__ validation (tensor_network.py:59-128) + SimpleUpdate.__init__
 * (simple_update.py:20-93).  `stream` is a cudaStream_t (NULL = the library creates its own). */
int tnsu_create(tnsu_engine **out, int n, int m, const int64_t *smat, int d, const int32_t *edge_dims,
                int d_max, int dtype, int batch, int device, void *stream);
int tnsu_destroy(tnsu_engine *e);

/* Shape queries (bond dimensions change while bonds grow towards d_max, simple_update.py:278-283). */
int tnsu_edge_dims(const tnsu_engine *e, int32_t *dims_out /* [m] */);
int64_t tnsu_tensor_elems(const tnsu_engine *e, int t);     /* current element count of tensor t */
int tnsu_num_levels(const tnsu_engine *e);                  /* dependency levels per sweep */
int tnsu_edge_levels(const tnsu_engine *e, int32_t *level_out /* [m] */);

/* State upload / download.  `host` holds `batch` tensors back to back, each in the engine dtype
 * (float64 or complex128) with the CURRENT shape (pageable or page-locked memory; the copy is complete on return).
 * The download entry points synchronise, and — like tnsu_sweep_error, tnsu_synchronize and tnsu_run — return
 * TNSU_E_NONFINITE (once) when an update since the last check met NaN / Inf: the reference raises ValueError from
 * scipy's finite check inside simple_update at that point (simple_update.py:243-250, :404).
 * Replaces: the list entries TensorNetwork.tensors[t] / .weights[e] (tensor_network.py:119-120). */
int tnsu_set_tensor(tnsu_engine *e, int t, const void *host, int64_t elems_per_point);
int tnsu_get_tensor(tnsu_engine *e, int t, void *host, int64_t elems_per_point);
int tnsu_set_weights(tnsu_engine *e, int edge, const double *host, int len_per_point);
int tnsu_get_weights(tnsu_engine *e, int edge, double *host, int len_per_point);

/* Two-site gates exp(-dt H_e) for time-step slot `slot`: complex128 [batch][m][d*d][d*d], computed by
 * the caller exactly as simple_update.py:420-473 does (numpy kron + scipy expm).  n_slots is fixed
 * by the first call to tnsu_reserve_gate_slots. */
int tnsu_reserve_gate_slots(tnsu_engine *e, int n_slots);
int tnsu_set_gates(tnsu_engine *e, int slot, const void *gates_c128);
/* Per-edge Hamiltonians for energy_per_site: complex128 [batch][m][d*d][d*d]
 * (simple_update.py:631-632). */
int tnsu_set_hamiltonians(tnsu_engine *e, const void *ham_c128);

/* n_sweeps Simple-Update sweeps over all edges with the gates of `slot`, every point.
 * Replaces: SimpleUpdate.simple_update() (simple_update.py:208-296).  After the call,
 * tnsu_sweep_error gives max_e ||lambda_e(new) - lambda_e(old)||_2 of the LAST sweep per point
 * (compute_convergence_error, simple_update.py:196-206). */
int tnsu_sweep(tnsu_engine *e, int slot, int n_sweeps);
int tnsu_sweep_error(tnsu_engine *e, double *err_out /* [batch] */);

/* Energy per site per point (simple_update.py:615-637). */
int tnsu_energy(tnsu_engine *e, double *energy_out /* [batch] */);
/* Two-site RDM of `edge`, trace-normalised, complex128 [batch][d][d][d][d] (simple_update.py:525-593). */
int tnsu_pair_rdm(tnsu_engine *e, int edge, void *rdm_c128);
/* One-site RDM of tensor t, complex128 [batch][d][d] (simple_update.py:498-523). */
int tnsu_site_rdm(tnsu_engine *e, int t, void *rdm_c128);
/* sum_e <O> / n  with O complex128 [d][d][d][d]  (pair_expectation_per_site, simple_update.py:639-649) and
 * sum_t tr(rho_t O) / n with O complex128 [d][d] (expectation_per_site, :651-661); complex128 out [batch]
 * (the caller takes the real part as the reference does). */
int tnsu_pair_expectation_per_site(tnsu_engine *e, const void *op_c128, void *out_c128);
int tnsu_expectation_per_site(tnsu_engine *e, const void *op_c128, void *out_c128);

/* Whole SimpleUpdate.run() for every point (simple_update.py:107-194): slots 0..n_dts-1 are the dt
 * schedule (gates must be set for each), `is_last_value[s]` != 0 where dts[s] == dts[-1] (the
 * reference compares VALUES, :176).  Per-point outputs; logs are [batch][log_cap] with
 * n_checks[b] valid entries (one per convergence check, i.e. every 2nd sweep of a dt, :132); final_slot[b] is the
 * last time-step slot the point entered (the reference leaves self.dt = dts[final_slot], :119). */
int tnsu_run(tnsu_engine *e, int n_dts, const int32_t *is_last_value, int max_iterations, double tol,
             int log_cap, int32_t *n_checks, double *log_error, double *log_energy, int32_t *log_slot,
             int32_t *log_step, int32_t *converged, int32_t *total_sweeps, int32_t *final_slot);

/* Timing / stream plumbing for benchmarks. */
void *tnsu_stream(tnsu_engine *e);
int tnsu_synchronize(tnsu_engine *e);
/* Number of kernels launched through this handle since creation. */
int64_t tnsu_launch_count(const tnsu_engine *e);
/* Average device time (ms, CUDA events on the engine stream) of the edge-update kernels launched by
 * the most recent tnsu_sweep call, and their count. */
int tnsu_last_sweep_kernel_ms(tnsu_engine *e, double *total_ms, int32_t *n_launches);

/* Cumulative counters of this handle (synchronises): out[0] kernels launched, out[1] (item, side) bond matrices given to
 * the streaming CholeskyQR2 reduction, out[2] of those flagged ill-conditioned and re-done by the Householder kernels,
 * out[3] Jacobi iterations that hit their sweep limit (the SVD is then only as converged as the limit allowed; the
 * reference's LAPACK driver would raise LinAlgError), out[4] non-finite events reported, out[5] tnsu_run calls served by
 * the on-chip persistent kernel.  At most n_out entries are written. */
#define TNSU_N_STATS 6
int tnsu_stats(tnsu_engine *e, int64_t *out, int n_out);

#ifdef __cplusplus
}
#endif
#endif /* TNSU_B200_H */
