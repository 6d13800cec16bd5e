// engine.cu — host side of libtnsu_b200.so: the C ABI of include/tnsu_b200.h, the edge scheduler
// (dependency levels that preserve the reference's sequential edge order, SURVEY.md §7 "Order
// dependence"), the bond-growth shape plan, kernel-variant selection and the batched run loop
// (SimpleUpdate.run, reference src/tnsu/simple_update.py:107-194) with per-point state on the device.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/tnsu_b200.h"
#include "../../include/tnsu_b200_debug.h"
#include "edge_update.cuh"
#include "rdm.cuh"
#include "run_fused.cuh"

static thread_local std::string g_create_error;

#define SMEM_LIMIT (227 * 1024)

struct LevelCfg {
  int mm_t = 0;      // MMAX template (8/16/32)
  int ncols = 1;     // NCOLS template
  int nt = 32;       // threads per side CTA
  int lda = 0, ldv = 0, nr_max = 0;
  int smem_lq = 0, smem_svd = 0, smem_rc = 0;
  int svdw_precond = 0, svdw_ldt = 0, svdw_accv = 1, svdw_tb_off = 0;
  int lqc = 0, lqc_m = 0;  // streaming CholeskyQR2 K1/K3 (lq_chol.cuh) in front of the Householder kernels; largest M of the level
  int rc_mout = 0, rc_k = 0, rc_ncl = 0;  // tensor-core K3 (recompose_dmma.cuh): largest Mout, K and slice width; 0 = not used
  int lq_stream_grid = 0, lq_stage_off = 0, lq_stage_doubles = 0;  // persistent TMA-staged K1
  int cluster = 1;                                                   // CTAs per (item, side) in K1 / K3
  int svdw_np = 0, svdw_rows = 0, svdw_group_doubles = 0;  // register-resident warp SVD (0 = shared-memory kernel)
  int svd_vfree = 0;                                         // shared-memory SVD kernel without the accumulator (edge_update.cuh)
};

struct tnsu_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int n = 0, m = 0, d = 0, d_max = 0, dtype = 0, batch = 0, d4 = 0;
  size_t ssz = 8;
  std::vector<int64_t> smat;
  struct Edge {
    int t[2];
    int leg[2];
  };
  std::vector<Edge> edges;
  std::vector<std::vector<int>> t_edges;  // per tensor: edge id on each leg (leg = axis-1)
  std::vector<int> level;
  std::vector<std::vector<int>> level_edges;
  std::vector<int> dims;                  // current bond dims
  std::vector<long long> cap;             // per-tensor capacity (scalars) per point
  int dcap = 1;
  // device state
  std::vector<void *> tens;
  std::vector<void *> ybuf;               // reflector scratch, same geometry as tens
  void *small = nullptr;                  // per-item small matrices
  size_t small_bytes = 0;
  int ldm = 1;                            // leading dimension of every small matrix in the current plan
  int max_level_size = 1;
  double prof_ms[3] = {0, 0, 0};
  bool profile = false;
  cudaEvent_t pev[4] = {nullptr, nullptr, nullptr, nullptr};
  double *weights = nullptr, *tscale = nullptr, *edge_err = nullptr, *energy = nullptr;
  double *tnorm2 = nullptr;               // K3 column slices: partial norms and arrival counters per (point, tensor)
  unsigned *tticket = nullptr;
  void *gates = nullptr, *hams = nullptr, *op_buf = nullptr;
  int n_slots = 0;
  std::vector<char> slot_set;
  bool hams_set = false;
  cplx *pair_val = nullptr, *cval = nullptr, *rdm_buf = nullptr, *site_buf = nullptr;
  EdgeDesc *d_descs = nullptr;
  std::vector<EdgeDesc> h_descs;
  std::vector<int> post_dims;
  bool plan_valid = false, plan_steady = false;
  bool lq_stream = false;                 // TNSU_LQ_STREAM=1: persistent TMA-staged K1 instead of one-shot CTAs (measured 4 % slower)
  int num_sms = 148;
  int lq_stream_min_n = 256;              // TNSU_LQ_MIN_N: bond matrices at least this wide stream at any batch size
  bool force_unrolled_lq = false;         // TNSU_LQ=unrolled: keep the fully unrolled K1 (A/B comparisons)
  bool svd_precond = true;                // TNSU_SVD_PRECOND=0: plain Jacobi in the warp SVD (A/B comparisons)
  bool use_graphs = false;                // TNSU_GRAPH=1: replay the run-loop iteration as a CUDA graph (measured: no faster)
  int force_ncols = 0;                    // TNSU_NCOLS=1|2: columns per thread in K1/K3 (experiments)
  bool svd_accumulate_v = false;          // TNSU_SVD_V=accumulate: warp SVD with the right rotations accumulated in registers (A/B comparisons)
  bool force_householder = false;         // TNSU_LQ=householder: no streaming CholeskyQR2 K1/K3 in front of the Householder kernels (A/B comparisons)
  int *lq_stats = nullptr;                // device counter: (item, side)s factored by the Householder fallback since the last poll
  long long lqc_attempted = 0;            // (item, side)s given to the streaming K1 since the last poll (host count)
  long long lqc_suspend_until = -1;       // run loop: no streaming K1/K3 before this global sweep index (most items were flagged)
  long long run_sweep_index = 0;          // global sweep index inside tnsu_run (for the suspension window)
  bool lqc_cta_kernel = false;            // TNSU_LQ=cta: two-warp-CTA streaming K1 instead of the one-warp-per-matrix kernel (A/B comparisons)
  bool lqc_force_fallback = false;        // TNSU_LQ=fallback: streaming kernel flags everything, the Householder fallback does all the work (tests)
  int *lq_flags = nullptr;                // [batch * max_level_size * 2] per (item, side): 1 = Householder kernels must process it
  bool force_simple_rc = false;           // TNSU_RC=simple: register-column K3 (recompose_kernel) instead of the tensor-core one (A/B comparisons)
  bool force_simple_rdm = false;          // TNSU_RDM=simple: shared-memory-staged pair RDM kernel instead of the DMMA one (A/B comparisons)
  bool force_smem_svd = false;            // TNSU_SVD=smem: keep the shared-memory Jacobi kernel (A/B comparisons)
  std::vector<LevelCfg> cfgs;
  int *d_level_edges = nullptr;
  int *d_level_off = nullptr;             // [n_levels + 1] (run_fused.cuh)
  int *d_toff = nullptr;                  // [n] shared-memory offsets of the tensors in the on-chip run kernel
  bool use_fused = true;                  // TNSU_FUSED=0: never use the on-chip persistent run kernel (A/B comparisons)
  bool force_fused = false;               // TNSU_FUSED=force: use it for every batch size it fits (A/B comparisons)
  std::vector<int> level_off;
  int *d_all_edges = nullptr;
  SiteDesc *d_site = nullptr;
  // run-loop state
  int *r_slot = nullptr, *r_iter = nullptr, *r_step = nullptr, *r_active = nullptr, *r_conv = nullptr,
      *r_nchecks = nullptr, *r_log_slot = nullptr, *r_log_step = nullptr, *r_nactive = nullptr, *r_islast = nullptr;
  double *r_log_err = nullptr, *r_log_energy = nullptr;
  int r_log_cap = 0;
  int *h_nactive = nullptr;  // pinned
  int *d_point_list = nullptr;  // tnsu_run: the points still active, in point order (compacted at the polls)
  int n_listed = -1;            // its length; -1: launches cover the whole batch
  bool compact_active = true;   // TNSU_COMPACT=0: keep whole-batch launches with the active mask (A/B comparisons)
  // failure counters (edge_update.cuh: EdgeLaunch::status) and cumulative statistics (tnsu_stats)
  int *d_status = nullptr;   // [2] device
  int *h_status = nullptr;   // [2] pinned
  int svd_max_sweeps = 1 << 20;  // TNSU_SVD_MAX_SWEEPS: lowers the kernels' Jacobi sweep limit (tests of the non-convergence counter)
  long long fused_runs = 0;  // tnsu_run calls served by the on-chip persistent kernel (run_fused.cuh)
  long long lqc_attempted_total = 0, lqc_flagged_total = 0, svd_unconverged_total = 0, nonfinite_events = 0;
  // debug buffers
  double *dbg_sigma = nullptr;
  int *dbg_info = nullptr;
  // bookkeeping
  std::string err;
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int last_n = 0;
  bool timed = false;
};

static int fail(tnsu_engine *e, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e)
    e->err = buf;
  else
    g_create_error = buf;
  return code;
}

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t _s = (call);                                                                         \
    if (_s != cudaSuccess) return fail(e, TNSU_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(_s)); \
  } while (0)

// Every ABI entry point runs on the engine's device and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  cudaError_t status = cudaSuccess;
  explicit DeviceGuard(int device) {
    status = cudaGetDevice(&prev);
    if (status != cudaSuccess) {
      prev = -1;
      return;
    }
    if (prev != device) status = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};
#define ON_DEVICE(e)                   \
  DeviceGuard _guard((e)->device);     \
  CK(_guard.status)

// Read the device failure counters (synchronises the stream).  A non-finite bond is reported ONCE, as the reference
// raises once from scipy's check_finite inside simple_update (simple_update.py:243-250, :404).
static int poll_status(tnsu_engine *e) {
  CK(cudaMemcpyAsync(e->h_status, e->d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const int bad = e->h_status[0], slow = e->h_status[1];
  if (bad || slow) CK(cudaMemsetAsync(e->d_status, 0, 2 * sizeof(int), e->stream));
  e->svd_unconverged_total += slow;
  if (bad) {
    ++e->nonfinite_events;
    return fail(e, TNSU_E_NONFINITE,
                "a bond update produced singular values that are not finite and positive: NaN / Inf in the tensors, "
                "weights or gates");
  }
  return TNSU_OK;
}

static int max_threads(bool cplx_, int mm, int nc) {
  int doubles = mm * nc * (cplx_ ? 2 : 1);
  return doubles > 48 ? 256 : 512;
}

// ---------------------------------------------------------------------------------------------
// shape plan: simulate one sweep in the reference's edge order from the current bond dims
// ---------------------------------------------------------------------------------------------
static int plan_sweep(tnsu_engine *e) {
  std::vector<int> cur = e->dims;
  e->h_descs.assign(e->m, EdgeDesc{});
  for (int ek = 0; ek < e->m; ++ek) {
    EdgeDesc &ed = e->h_descs[ek];
    ed.edge = ek;
    ed.d = e->d;
    ed.Dk = cur[ek];
    for (int sd_i = 0; sd_i < 2; ++sd_i) {
      SideDesc &sd = ed.s[sd_i];
      int t = e->edges[ek].t[sd_i];
      const std::vector<int> &legs = e->t_edges[t];
      sd.tensor = t;
      sd.base = e->tens[t];
      sd.ybase = e->ybuf[t];
      sd.point_stride = e->cap[t];
      sd.nlegs = (int)legs.size();
      sd.bond_leg = e->edges[ek].leg[sd_i];
      long long nprod = 1;
      for (int l = 0; l < sd.nlegs; ++l) {
        sd.dims[l] = cur[legs[l]];
        sd.leg_edge[l] = legs[l];
        if (l != sd.bond_leg) nprod *= sd.dims[l];
      }
      if (nprod > (1 << 20)) return fail(e, TNSU_E_UNSUPPORTED, "bond matrix of tensor %d has %lld columns", t, nprod);
      sd.M = e->d * ed.Dk;
      sd.N = (int)nprod;
      sd.K = std::min(sd.M, sd.N);
    }
    ed.ni = ed.s[0].K * e->d;
    ed.nj = e->d * ed.s[1].K;
    ed.Dnew = std::min(e->d_max, std::min(ed.ni, ed.nj));
    for (int sd_i = 0; sd_i < 2; ++sd_i) {
      SideDesc &sd = ed.s[sd_i];
      sd.Mout = e->d * ed.Dnew;
      long long si = 1, so = 1;
      for (int l = sd.nlegs - 1; l >= 0; --l) {
        sd.in_stride[l] = si;
        sd.out_stride[l] = so;
        si *= sd.dims[l];
        so *= (l == sd.bond_leg) ? ed.Dnew : sd.dims[l];
      }
      sd.in_stride_phys = si;
      sd.out_stride_phys = so;
      if (si * e->d > e->cap[sd.tensor] || so * e->d > e->cap[sd.tensor])
        return fail(e, TNSU_E_STATE, "tensor %d outgrew its capacity", sd.tensor);
    }
    if (ed.Dnew > e->dcap) return fail(e, TNSU_E_STATE, "bond %d outgrew the weight capacity", ek);
    cur[ek] = ed.Dnew;
  }
  e->post_dims = cur;
  e->plan_steady = (cur == e->dims);

  // per-level kernel configuration
  e->cfgs.assign(e->level_edges.size(), LevelCfg{});
  const bool cx = e->dtype == TNSU_C128;
  int ldm_all = 1;
  for (int ek = 0; ek < e->m; ++ek)
    for (int s = 0; s < 2; ++s) {
      const SideDesc &sd = e->h_descs[ek].s[s];
      ldm_all = std::max(ldm_all, std::max(sd.M, std::max(sd.Mout, sd.K)));
    }
  if (ldm_all > TNSU_MAX_M)
    return fail(e, TNSU_E_UNSUPPORTED, "bond matrix has %d rows (d*D); this build holds at most %d", ldm_all, TNSU_MAX_M);
  e->ldm = ldm_all | 1;
  const int ssz = (int)e->ssz;
  for (size_t lv = 0; lv < e->level_edges.size(); ++lv) {
    int maxM = 1, maxN = 1, maxnr = 1, maxnc = 1;
    for (int ek : e->level_edges[lv]) {
      const EdgeDesc &ed = e->h_descs[ek];
      for (int s = 0; s < 2; ++s) {
        maxM = std::max(maxM, std::max(ed.s[s].M, std::max(ed.s[s].Mout, ed.s[s].K)));
        maxN = std::max(maxN, ed.s[s].N);
      }
      maxnr = std::max(maxnr, std::max(ed.ni, ed.nj));
      maxnc = std::max(maxnc, std::min(ed.ni, ed.nj));
    }
    LevelCfg c;
    c.mm_t = maxM <= 8 ? 8 : (maxM <= 16 ? 16 : 32);
    if (maxnr > TNSU_MAX_SVD)
      return fail(e, TNSU_E_UNSUPPORTED, "theta has %d rows; this build holds at most %d", maxnr, TNSU_MAX_SVD);
    bool ok = false;
    const int order2[2] = {2, 1}, order1[2] = {1, 2};
    const int *ncs = (maxN > 128) ? order2 : order1;  // two columns per thread once the matrix is wide
    if (e->force_ncols == 1) ncs = order1;
    if (e->force_ncols == 2) ncs = order2;
    for (int o = 0; o < 2 && !ok; ++o) {
      const int nc = ncs[o];
      if (cx && c.mm_t == 32 && nc == 2) continue;  // not instantiated: 256 data registers per thread
      int nt = ((maxN + nc - 1) / nc + 31) & ~31;
      if (nt > max_threads(cx, c.mm_t, nc) || nt > TNSU_MAX_WARPS * 32) continue;
      c.ncols = nc;
      c.nt = nt;
      ok = true;
    }
    if (!ok && !(e->force_unrolled_lq && !cx)) {
      // wide bond matrices: split the columns over a thread-block cluster of 2, 4 or 8 CTAs (K1) / column slices (K3)
      for (int cl = 2; cl <= 8 && !ok; cl *= 2) {
        const int ncl = (maxN + cl - 1) / cl;
        for (int o = 0; o < 2 && !ok; ++o) {
          const int nc = order2[o];
          if (cx && c.mm_t == 32 && nc == 2) continue;  // not instantiated
          const int nt = ((ncl + nc - 1) / nc + 31) & ~31;
          if (nt > max_threads(cx, c.mm_t, nc) || nt > TNSU_MAX_WARPS * 32) continue;
          c.ncols = nc;
          c.nt = nt;
          c.cluster = cl;
          ok = true;
        }
      }
    }
    if (!ok)
      return fail(e, TNSU_E_UNSUPPORTED,
                  "no kernel variant holds a %d x %d bond matrix (%s) in registers; reduce the bond dimension",
                  maxM, maxN, cx ? "complex128" : "float64");
    c.lda = maxnr | 1;
    c.nr_max = maxnr;
    c.ldv = maxnc | 1;
    const int ldm2 = e->ldm * e->ldm;
    c.smem_lq = TNSU_MAX_LEGS * e->dcap * 8 + c.mm_t * 8 + (2 * TNSU_MAX_WARPS * c.mm_t + 2 * c.mm_t + 4 * ldm2) * ssz +
                c.mm_t * 8 + 64;
    if (!cx) c.smem_lq += (c.nt / 32) * c.mm_t * 33 * 8;  // per-warp transpose buffers of lq_rolled_kernel
    if (c.cluster > 1) {
      c.lq_stage_off = (c.smem_lq + 15) & ~15;  // the cluster's partial Gram rows [2][MMAX]
      c.smem_lq = c.lq_stage_off + 2 * c.mm_t * ssz + 16;
    }
    if (!cx && e->lq_stream && c.cluster == 1) {
      // streaming K1: one staging buffer for the largest site tensor of the level; two CTAs must fit one SM
      long long stage = 0;
      for (int ek : e->level_edges[lv])
        for (int s2 = 0; s2 < 2; ++s2) stage = std::max(stage, e->h_descs[ek].s[s2].in_stride_phys * (long long)e->d);
      stage = (stage + 1) & ~1LL;
      const int off = (c.smem_lq + 15) & ~15;
      const long long total = off + stage * 8 + 16;
      if (2 * (total + 1024) <= SMEM_LIMIT) {
        c.lq_stage_off = off;
        c.lq_stage_doubles = (int)stage;
        c.smem_lq = (int)total;
        c.lq_stream_grid = 2 * e->num_sms;
      }
    }
    c.svd_vfree = e->svd_accumulate_v ? 0 : 1;
    c.smem_svd = (2 * ldm2 + e->d4 + c.lda * c.ldv +
                  (c.svd_vfree ? c.lda * c.ldv + e->dcap * c.lda : (c.ldv * c.ldv <= 2 * ldm2 ? 0 : c.ldv * c.ldv))) * ssz +
                 2 * TNSU_MAX_SVD * 8 + TNSU_MAX_SVD * 4 + 64;
    if (c.svd_vfree && c.smem_svd > SMEM_LIMIT) {  // the accumulating variant needs less shared memory for very large thetas
      c.svd_vfree = 0;
      c.smem_svd = (2 * ldm2 + e->d4 + c.lda * c.ldv + (c.ldv * c.ldv <= 2 * ldm2 ? 0 : c.ldv * c.ldv)) * ssz +
                   2 * TNSU_MAX_SVD * 8 + TNSU_MAX_SVD * 4 + 64;
    }
    c.smem_rc = TNSU_MAX_LEGS * e->dcap * 8 + c.mm_t * 8 + 3 * ldm2 * ssz + TNSU_MAX_WARPS * 8 + 64;
    // real thetas of at most 32 x 32 take the register-resident warp kernel (svd_warp.cuh)
    if (!cx && maxnr <= 32 && !e->force_smem_svd) {
      const int need = (maxnr + 1) / 2;
      const int rt[6] = {2, 4, 6, 8, 12, 16};
      for (int k = 0; k < 6; ++k)
        if (rt[k] >= need) {
          c.svdw_rows = rt[k];
          break;
        }
      c.svdw_np = std::max(1, (maxnc + 1) / 2);
      c.svdw_group_doubles = 2 * ldm2 + e->d4 + e->dcap + 2 * c.svdw_rows * 2 * c.svdw_np;
      if (e->svd_precond && maxnr >= 8) {
        // one column of A' = theta^T (nj x ni) and of Q^T (nj x nj) per lane: the group spans max(ni, nj) lanes
        c.svdw_precond = 1;
        c.svdw_np = std::max(1, (maxnr + 1) / 2);
        c.svdw_ldt = std::max(2 * c.svdw_np, 2 * c.svdw_rows) | 1;
        c.svdw_group_doubles = 2 * ldm2 + e->d4 + e->dcap + 2 * c.svdw_rows * c.svdw_ldt + 2 * c.svdw_rows + 2;
        if (!e->svd_accumulate_v) {
          // V-free iteration: theta (2R x 2NP, column-major) stays in place; the half-size transpose buffer and the
          // pivot column go on the dead L_i / L_j when they fit, else behind theta
          c.svdw_accv = 0;
          const int tb = c.svdw_rows * c.svdw_ldt + 2 * c.svdw_rows + 2;
          const int th = 2 * c.svdw_rows * 2 * c.svdw_rows;  // read without bounds at the end: full 2R x 2R, zero padded
          if (tb <= 2 * ldm2) {
            c.svdw_tb_off = 0;
            c.svdw_group_doubles = 2 * ldm2 + e->d4 + e->dcap + th;
          } else {
            c.svdw_tb_off = 2 * ldm2 + e->d4 + e->dcap + th;
            c.svdw_group_doubles = c.svdw_tb_off + tb;
          }
        }
      }
      c.svdw_group_doubles = (c.svdw_group_doubles + 1) & ~1;
    }
    if (!e->force_householder && !e->force_simple_rc && !e->lq_stream && !e->force_unrolled_lq && e->dcap <= 32 &&
        (cx ? 2 * maxM <= 32 : c.cluster == 1) &&
        (maxN >= e->lq_stream_min_n || 2LL * e->batch * (long long)e->level_edges[lv].size() >= 16LL * e->num_sms)) {
      // streaming K1/K3: every side of the level must have M <= N, an unchanged bond dimension (the recompose runs in
      // place) and tensors below 2 GiB (32-bit byte offsets).  Narrow bond matrices (N < 256: the D <= 4 networks of the
      // sweep metric) stay with the Householder kernels while a launch is latency bound — two Cholesky chains plus two
      // more launches per level cost more than they save (measured at D = 4: 608 vs 570 us per sweep at 256 points) —
      // and stream as soon as the launch fills the GPU (one wave of 16 warps per SM): 1.55 vs 1.86 ms per sweep at 4096
      // points, 10.3 vs 12.7 ms at 32768
      bool ok_lqc = true;
      int mm = 1;
      for (int ek : e->level_edges[lv]) {
        const EdgeDesc &ed = e->h_descs[ek];
        if (ed.Dnew != ed.Dk) ok_lqc = false;
        for (int s2 = 0; s2 < 2; ++s2) {
          const SideDesc &sd2 = ed.s[s2];
          if (sd2.M > sd2.N || sd2.Mout != sd2.M) ok_lqc = false;
          if ((long long)sd2.M * sd2.N * ssz >= (1LL << 31)) ok_lqc = false;
          for (int l = 0; l < sd2.nlegs; ++l)
            if (sd2.in_stride[l] != sd2.out_stride[l]) ok_lqc = false;
          if (sd2.in_stride_phys != sd2.out_stride_phys) ok_lqc = false;
          mm = std::max(mm, sd2.M);
        }
      }
      if (ok_lqc) {
        c.lqc = 1;
        c.lqc_m = mm;
      }
    }
    if (!cx && !e->force_simple_rc && e->dcap <= 32) {
      int mo = 1, kk = 1;
      for (int ek : e->level_edges[lv])
        for (int s2 = 0; s2 < 2; ++s2) {
          mo = std::max(mo, e->h_descs[ek].s[s2].Mout);
          kk = std::max(kk, e->h_descs[ek].s[s2].K);
        }
      c.rc_mout = mo;
      c.rc_k = kk;
      c.rc_ncl = (maxN + c.cluster - 1) / c.cluster;
    }
    if (std::max(c.smem_lq, std::max(c.smem_svd, c.smem_rc)) > SMEM_LIMIT)
      return fail(e, TNSU_E_UNSUPPORTED, "edge update needs %d bytes of shared memory", std::max(c.smem_lq, c.smem_svd));
    e->cfgs[lv] = c;
  }
  {
    size_t need = (size_t)e->batch * e->max_level_size * SM_COUNT * e->ldm * e->ldm * e->ssz;
    if (need > e->small_bytes) {
      cudaStreamSynchronize(e->stream);
      if (e->small) cudaFree(e->small);
      e->small = nullptr;
      cudaError_t st2 = cudaMalloc(&e->small, need);
      if (st2 != cudaSuccess) return fail(e, TNSU_E_CUDA, "scratch allocation of %zu bytes: %s", need, cudaGetErrorString(st2));
      e->small_bytes = need;
    }
    if (e->lq_flags == nullptr) {
      cudaError_t st3 = cudaMalloc(&e->lq_flags, sizeof(int) * 2 * (size_t)e->batch * e->max_level_size);
      if (st3 == cudaSuccess) st3 = cudaMalloc(&e->lq_stats, sizeof(int));
      if (st3 == cudaSuccess) st3 = cudaMemsetAsync(e->lq_stats, 0, sizeof(int), e->stream);
      if (st3 != cudaSuccess) return fail(e, TNSU_E_CUDA, "flag allocation: %s", cudaGetErrorString(st3));
    }
  }
  cudaError_t st = cudaMemcpyAsync(e->d_descs, e->h_descs.data(), sizeof(EdgeDesc) * e->m, cudaMemcpyHostToDevice,
                                   e->stream);
  if (st != cudaSuccess) return fail(e, TNSU_E_CUDA, "descriptor upload: %s", cudaGetErrorString(st));
  // h_descs is pageable: the async copy is staged before returning, safe to reuse
  e->plan_valid = true;
  return TNSU_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------------------------
// the kernel variants are instantiated in edge_inst.cu (one translation unit per scalar type and MMAX,
// compiled in parallel); here they are only declared
#define TNSU_EXTERN_VARIANT(S, MM, NC)                                                                   \
  extern template cudaError_t launch_lq_variant<S, MM, NC>(const EdgeLaunch<S> &, int, cudaStream_t);    \
  extern template cudaError_t launch_rc_variant<S, MM, NC>(const EdgeLaunch<S> &, int, cudaStream_t);
TNSU_EXTERN_VARIANT(double, 8, 1)
TNSU_EXTERN_VARIANT(double, 8, 2)
TNSU_EXTERN_VARIANT(double, 16, 1)
TNSU_EXTERN_VARIANT(double, 16, 2)
TNSU_EXTERN_VARIANT(double, 32, 1)
TNSU_EXTERN_VARIANT(double, 32, 2)
TNSU_EXTERN_VARIANT(cplx, 8, 1)
TNSU_EXTERN_VARIANT(cplx, 8, 2)
TNSU_EXTERN_VARIANT(cplx, 16, 1)
TNSU_EXTERN_VARIANT(cplx, 16, 2)
TNSU_EXTERN_VARIANT(cplx, 32, 1)
extern template cudaError_t launch_svd<double>(const EdgeLaunch<double> &, int, cudaStream_t);
extern template cudaError_t launch_svd<cplx>(const EdgeLaunch<cplx> &, int, cudaStream_t);

template <typename S>
static cudaError_t launch_side_kernel(bool recompose, const LevelCfg &c, const EdgeLaunch<S> &p, int n_items,
                                      cudaStream_t st) {
#define VAR(MM, NC)                         \
  if (c.mm_t == MM && c.ncols == NC)        \
    return recompose ? launch_rc_variant<S, MM, NC>(p, n_items, st) : launch_lq_variant<S, MM, NC>(p, n_items, st);
  VAR(8, 1)
  VAR(8, 2)
  VAR(16, 1)
  VAR(16, 2)
  VAR(32, 1)
  if constexpr (sizeof(S) == 8) {
    VAR(32, 2)
  }
#undef VAR
  return cudaErrorInvalidConfiguration;
}

cudaError_t launch_svd_warp_dispatch(const EdgeLaunch<double> &p, int n_items, cudaStream_t st);  // svdw_inst.cu
cudaError_t launch_lq_chol_dispatch(const EdgeLaunch<double> &p, int n_items, int m_max, int *flags, cudaStream_t st);  // lqc_inst.cu
cudaError_t launch_rc_chol_dispatch(const EdgeLaunch<double> &p, int n_items, int m_max, const int *flags, cudaStream_t st);
cudaError_t launch_lq_chol_cx_dispatch(const EdgeLaunch<cplx> &p, int n_items, int m_max, int *flags, cudaStream_t st);
cudaError_t launch_rc_chol_cx_dispatch(const EdgeLaunch<cplx> &p, int n_items, int m_max, const int *flags, cudaStream_t st);
static cudaError_t launch_k1c(const EdgeLaunch<double> &p, int n, int m, int *f, cudaStream_t st) { return launch_lq_chol_dispatch(p, n, m, f, st); }
static cudaError_t launch_k1c(const EdgeLaunch<cplx> &p, int n, int m, int *f, cudaStream_t st) { return launch_lq_chol_cx_dispatch(p, n, m, f, st); }
static cudaError_t launch_k3c(const EdgeLaunch<double> &p, int n, int m, const int *f, cudaStream_t st) { return launch_rc_chol_dispatch(p, n, m, f, st); }
static cudaError_t launch_k3c(const EdgeLaunch<cplx> &p, int n, int m, const int *f, cudaStream_t st) { return launch_rc_chol_cx_dispatch(p, n, m, f, st); }
cudaError_t launch_recompose_dmma_dispatch(const EdgeLaunch<double> &p, int n_items, int mout_max, int k_max, int ncl_cap,
                                           cudaStream_t st);  // rc_inst.cu

template <typename S>
static cudaError_t launch_k3(const LevelCfg &c, const EdgeLaunch<S> &p, int n_items, cudaStream_t st) {
  if constexpr (sizeof(S) == 8) {
    if (c.rc_mout > 0) return launch_recompose_dmma_dispatch(p, n_items, c.rc_mout, c.rc_k, c.rc_ncl, st);
  }
  return launch_side_kernel<S>(true, c, p, n_items, st);
}

template <typename S>
static cudaError_t launch_k2(const LevelCfg &c, const EdgeLaunch<S> &p, int n_items, cudaStream_t st) {
  if constexpr (sizeof(S) == 8) {
    if (c.svdw_np > 0) return launch_svd_warp_dispatch(p, n_items, st);
  }
  return launch_svd<S>(p, n_items, st);
}

template <typename S>
static int run_levels(tnsu_engine *e, int fixed_slot, const int *point_slot, const int *active, int only_edge,
                      int dbg_point) {
  for (size_t lv = 0; lv < e->level_edges.size(); ++lv) {
    const LevelCfg &c = e->cfgs[lv];
    EdgeLaunch<S> p{};
    p.descs = e->d_descs;
    p.level_edges = e->d_level_edges + e->level_off[lv];
    p.n_level_edges = (int)e->level_edges[lv].size();
    if (only_edge >= 0) {
      if (e->level[only_edge] != (int)lv) continue;
      p.level_edges = e->d_all_edges + only_edge;
      p.n_level_edges = 1;
    }
    p.point_list = (active != nullptr && e->n_listed >= 0) ? e->d_point_list : nullptr;
    p.batch = p.point_list != nullptr ? e->n_listed : e->batch;
    p.m_edges = e->m;
    p.dcap = e->dcap;
    p.d4 = e->d4;
    p.n_tensors = e->n;
    p.weights = e->weights;
    p.tscale = e->tscale;
    p.gates = (const S *)e->gates;
    p.gate_slot_stride = (long long)e->batch * e->m * e->d4;
    p.point_slot = point_slot;
    p.fixed_slot = fixed_slot;
    p.active = active;
    p.edge_err = e->edge_err;
    p.small = (S *)e->small;
    p.ldm = e->ldm;
    p.lda = c.lda;
    p.ldv = c.ldv;
    p.nr_max = c.nr_max;
    p.nt = c.nt;
    p.smem_lq = c.smem_lq;
    p.smem_svd = c.smem_svd;
    p.smem_rc = c.smem_rc;
    p.lq_rolled = (!e->force_unrolled_lq && e->dtype == TNSU_F64) ? 1 : 0;
    p.cluster = c.cluster;
    p.tnorm2 = e->tnorm2;
    p.tticket = e->tticket;
    p.lq_stream_grid = c.lq_stream_grid;
    p.lq_stage_off = c.lq_stage_off;
    p.lq_stage_doubles = c.lq_stage_doubles;
    p.svdw_np = c.svdw_np;
    p.svdw_rows = c.svdw_rows;
    p.svdw_group_doubles = c.svdw_group_doubles;
    p.svdw_precond = c.svdw_precond;
    p.svdw_ldt = c.svdw_ldt;
    p.lq_flags = nullptr;
    p.lq_stats = nullptr;
    p.lqc_force_fallback = e->lqc_force_fallback ? 1 : (e->lqc_cta_kernel ? 2 : 0);
    p.svdw_accv = c.svdw_accv;
    p.svd_vfree = c.svd_vfree;
    p.svd_precond = e->svd_precond ? 1 : 0;
    p.svdw_tb_off = c.svdw_tb_off;
    p.status = e->d_status;
    p.svd_max_sweeps = e->svd_max_sweeps;
    p.dbg_info = nullptr;
    if (dbg_point >= 0) {
      p.dbg_sigma = e->dbg_sigma;
      p.dbg_info = e->dbg_info;
      p.dbg_point = dbg_point;
    }
    const int n_items = p.batch * p.n_level_edges;
    cudaError_t st = cudaSuccess;
    if (e->profile) cudaEventRecord(e->pev[0], e->stream);
    int extra_launches = 0;
    if (c.lqc && only_edge < 0 && e->run_sweep_index >= e->lqc_suspend_until) {
      // streaming CholeskyQR2 first; it flags the (item, side)s it cannot take and the Householder kernels below
      // then process exactly those (every other CTA of theirs returns at once)
      p.lq_flags = e->lq_flags;
      p.lq_stats = e->lq_stats;
      e->lqc_attempted += 2LL * n_items;
      e->lqc_attempted_total += 2LL * n_items;
      st = launch_k1c(p, n_items, c.lqc_m, e->lq_flags, e->stream);
      ++extra_launches;
    }
    if (st == cudaSuccess) st = launch_side_kernel<S>(false, c, p, n_items, e->stream);
    if (e->profile) cudaEventRecord(e->pev[1], e->stream);
    if (st == cudaSuccess) st = launch_k2<S>(c, p, n_items, e->stream);
    if (e->profile) cudaEventRecord(e->pev[2], e->stream);
    if (st == cudaSuccess && p.lq_flags != nullptr) {
      st = launch_k3c(p, n_items, c.lqc_m, e->lq_flags, e->stream);
      ++extra_launches;
    }
    if (st == cudaSuccess) st = launch_k3<S>(c, p, n_items, e->stream);
    if (e->profile) {
      cudaEventRecord(e->pev[3], e->stream);
      cudaEventSynchronize(e->pev[3]);
      for (int k = 0; k < 3; ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e->pev[k], e->pev[k + 1]);
        e->prof_ms[k] += ms;
      }
    }
    if (st != cudaSuccess)
      return fail(e, TNSU_E_CUDA, "edge kernel launch (level %zu, MM=%d NC=%d nt=%d smem=%d/%d/%d): %s", lv, c.mm_t,
                  c.ncols, c.nt, c.smem_lq, c.smem_svd, c.smem_rc, cudaGetErrorString(st));
    e->launches += 3 + extra_launches;
    e->last_n += 3 + extra_launches;
  }
  return TNSU_OK;
}

static int one_sweep(tnsu_engine *e, int fixed_slot, const int *point_slot, const int *active) {
  if (!e->plan_valid) {
    int rc = plan_sweep(e);
    if (rc) return rc;
  }
  int rc = (e->dtype == TNSU_C128) ? run_levels<cplx>(e, fixed_slot, point_slot, active, -1, -1)
                                   : run_levels<double>(e, fixed_slot, point_slot, active, -1, -1);
  if (rc) return rc;
  if (!e->plan_steady) {
    e->dims = e->post_dims;
    e->plan_valid = false;
  }
  return TNSU_OK;
}

// ---------------------------------------------------------------------------------------------
// RDM launches
// ---------------------------------------------------------------------------------------------
static int pair_smem_bytes(const tnsu_engine *e, int M) {
  return (int)(2 * M * M * e->ssz + M * (RDM_CHUNK + 1) * e->ssz + RDM_CHUNK * 8 + RDM_CHUNK * 8 +
               TNSU_MAX_LEGS * e->dcap * 8 + e->dcap * 8 + e->d4 * 16 + 64 * 16);
}

template <typename S>
static int launch_pair(tnsu_engine *e, const int *edge_list, int n_list, const S *ops, long long op_pstride,
                       long long op_estride, cplx *pair_val, cplx *rdm_out, const int *iter_state, const int *active) {
  // the caller (ensure_static_descs) has put descriptors of the CURRENT shapes on the device
  int maxM = 1;
  for (int ek = 0; ek < e->m; ++ek) maxM = std::max(maxM, e->d * e->dims[ek]);
  PairRdmLaunch<S> p{};
  p.descs = e->d_descs;
  p.edge_list = edge_list;
  p.first_edge = (edge_list >= e->d_all_edges && edge_list < e->d_all_edges + e->m) ? (int)(edge_list - e->d_all_edges) : -1;
  p.n_list = n_list;
  p.point_list = (active != nullptr && e->n_listed >= 0) ? e->d_point_list : nullptr;
  p.batch = p.point_list != nullptr ? e->n_listed : e->batch;
  p.m_edges = e->m;
  p.dcap = e->dcap;
  p.d4 = e->d4;
  p.weights = e->weights;
  p.ops = ops;
  p.op_point_stride = op_pstride;
  p.op_edge_stride = op_estride;
  p.pair_val = pair_val;
  p.rdm_out = rdm_out;
  p.iter_state = iter_state;
  p.active = active;
  if constexpr (std::is_same<S, double>::value) {
    if (!e->force_simple_rdm && e->dcap <= PG_WL) {
      // real path: Gram matrices on the FP64 tensor cores, operands straight from HBM/L2 (rdm.cuh)
      const int MB = (maxM + 7) / 8;
      const int smem_d = pair_dmma_smem_bytes(MB, e->d4);
      const int grid = p.batch * n_list;
      switch (MB) {
        case 1: pair_rdm_dmma_kernel<1><<<grid, PG_THREADS, smem_d, e->stream>>>(p); break;
        case 2: pair_rdm_dmma_kernel<2><<<grid, PG_THREADS, smem_d, e->stream>>>(p); break;
        case 3: pair_rdm_dmma_kernel<3><<<grid, PG_THREADS, smem_d, e->stream>>>(p); break;
        default: pair_rdm_dmma_kernel<4><<<grid, PG_THREADS, smem_d, e->stream>>>(p); break;
      }
      CK(cudaGetLastError());
      e->launches++;
      return TNSU_OK;
    }
  }
  int smem = pair_smem_bytes(e, maxM);
  if (smem > SMEM_LIMIT) return fail(e, TNSU_E_UNSUPPORTED, "pair RDM needs %d bytes of shared memory", smem);
  CK(tnsu_opt_in_smem((const void *)pair_rdm_kernel<S>));
  pair_rdm_kernel<S><<<p.batch * n_list, RDM_THREADS, smem, e->stream>>>(p);
  CK(cudaGetLastError());
  e->launches++;
  return TNSU_OK;
}

// RDMs need descriptors of the CURRENT shapes for every edge (not the mid-sweep ones).  In steady state
// the sweep plan is exactly that; otherwise build a static plan with Dnew = Dk semantics.
static int ensure_static_descs(tnsu_engine *e) {
  if (!e->plan_valid) {
    int rc = plan_sweep(e);
    if (rc) return rc;
  }
  if (e->plan_steady) return TNSU_OK;
  // not steady: overwrite the device descriptors with "current shape" descriptors (input side only is read)
  std::vector<EdgeDesc> st = e->h_descs;
  for (int ek = 0; ek < e->m; ++ek) {
    EdgeDesc &ed = st[ek];
    ed.Dk = e->dims[ek];
    for (int s = 0; s < 2; ++s) {
      SideDesc &sd = ed.s[s];
      const std::vector<int> &legs = e->t_edges[sd.tensor];
      long long nprod = 1, si = 1;
      for (int l = 0; l < sd.nlegs; ++l) sd.dims[l] = e->dims[legs[l]];
      for (int l = sd.nlegs - 1; l >= 0; --l) {
        sd.in_stride[l] = si;
        si *= sd.dims[l];
        if (l != sd.bond_leg) nprod *= sd.dims[l];
      }
      sd.in_stride_phys = si;
      sd.M = e->d * ed.Dk;
      sd.N = (int)nprod;
    }
  }
  CK(cudaMemcpyAsync(e->d_descs, st.data(), sizeof(EdgeDesc) * e->m, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->plan_valid = false;  // the sweep plan on the device was overwritten; rebuild before the next sweep
  return TNSU_OK;
}

static int upload_site_descs(tnsu_engine *e) {
  std::vector<SiteDesc> sd(e->n);
  for (int t = 0; t < e->n; ++t) {
    sd[t].base = e->tens[t];
    sd[t].point_stride = e->cap[t];
    sd[t].nlegs = (int)e->t_edges[t].size();
    long long nv = 1;
    for (int l = 0; l < sd[t].nlegs; ++l) {
      sd[t].dims[l] = e->dims[e->t_edges[t][l]];
      sd[t].leg_edge[l] = e->t_edges[t][l];
      nv *= sd[t].dims[l];
    }
    sd[t].nvirt = nv;
  }
  CK(cudaMemcpyAsync(e->d_site, sd.data(), sizeof(SiteDesc) * e->n, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TNSU_OK;
}

// ---------------------------------------------------------------------------------------------
// run-loop state machine (simple_update.py:117-194), one thread per point
// ---------------------------------------------------------------------------------------------
__global__ void run_state_kernel(RunState s) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= s.batch || s.active[b] == 0) return;
  int i = s.iter[b];
  int slot = s.slot[b];
  s.step[b] += 1;
  bool advance = false;
  if (i % 2 == 0 && i > 0) {
    double err = 0.0;
    bool bad = false;
    for (int e = 0; e < s.m_edges; ++e) {
      double v = s.edge_err[(size_t)b * s.m_edges + e];
      if (v != v) bad = true;
      err = fmax(err, v);
    }
    if (bad) err = nan("");
    const int k = s.nchecks[b];
    if (k < s.log_cap) {
      s.log_err[(size_t)b * s.log_cap + k] = err;
      s.log_energy[(size_t)b * s.log_cap + k] = s.energy[b];
      s.log_slot[(size_t)b * s.log_cap + k] = slot;
      s.log_step[(size_t)b * s.log_cap + k] = s.step[b];
    }
    s.nchecks[b] = k + 1;
    if (err <= s.tol && s.is_last[slot]) {
      s.conv[b] = 1;
      s.active[b] = 0;
      return;
    }
    if (err <= s.tol) advance = true;
  }
  if (!advance) {
    i += 1;
    if (i >= s.max_iter)
      advance = true;
    else
      s.iter[b] = i;
  }
  if (advance) {
    slot += 1;
    s.iter[b] = 0;
    if (slot >= s.n_dts) {
      s.active[b] = 0;
      return;
    }
    s.slot[b] = slot;
  }
  atomicAdd(s.nactive, 1);
}

// Stable compaction of the active points into `list` (one block; the run loop calls it at a poll when points finished):
// the edge-update and pair-RDM launches that follow cover exactly the listed points, so a finished point costs nothing —
// with the `active` mask alone its warps return at once, but its CTAs still take part in every launch and leave the
// CTAs of the remaining points, whose kernels are latency bound, half empty.
__global__ void compact_active_kernel(const int *active, int n, int *list) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int start = 0; start < n; start += (int)blockDim.x) {
    const int i = start + (int)threadIdx.x;
    const bool a = i < n && active[i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, a);
    if (lane == 0) warp_tot[w] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int k = 0; k < w; ++k) off += warp_tot[k];
    if (a) list[off + __popc(bal & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += warp_tot[k];
      base += t;
    }
    __syncthreads();
  }
}

__global__ void fill_int_kernel(int *p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill_double_kernel(double *p, long long n, double v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// on-chip persistent run kernel (run_fused.cuh): eligibility, shared-memory layout, launch
// ---------------------------------------------------------------------------------------------
// Returns 1 when the rest of the run was handed to run_fused_kernel, 0 when this network does not qualify (the caller
// continues with the general launch path), or a negative error code.
static int try_run_fused(tnsu_engine *e, const RunState &st) {
  if (!e->use_fused || e->dtype != TNSU_F64 || !e->plan_valid || !e->plan_steady) return 0;
  int maxnr = 1, maxnc = 1, maxM = 1, maxN = 1, width = 1;
  for (int ek = 0; ek < e->m; ++ek) {
    const EdgeDesc &ed = e->h_descs[ek];
    if (ed.Dnew != ed.Dk) return 0;
    maxnr = std::max(maxnr, std::max(ed.ni, ed.nj));
    maxnc = std::max(maxnc, std::min(ed.ni, ed.nj));
    for (int s = 0; s < 2; ++s) {
      maxM = std::max(maxM, std::max(ed.s[s].M, ed.s[s].Mout));
      maxN = std::max(maxN, ed.s[s].N);
    }
  }
  if (maxnr > 32 || maxM > 32 || e->dcap > 32) return 0;
  for (auto &lv : e->level_edges) width = std::max(width, (int)lv.size());
  FusedLaunch p{};
  const int rt[6] = {2, 4, 6, 8, 12, 16};
  int rows = 16;
  for (int k = 0; k < 6; ++k)
    if (rt[k] >= (maxnr + 1) / 2) {
      rows = rt[k];
      break;
    }
  const int ldm2 = e->ldm * e->ldm;
  if (e->svd_precond && maxnr >= 8) {
    // the same geometry as plan_sweep gives svd_warp_kernel: pivoted-QR preconditioner, V-free iteration (measured in this
    // kernel: the accumulating variant's QR with its Q^T rows costs 46k cycles per 16 x 16 theta against 21k)
    p.svd_precond = 1;
    p.svd_np = std::max(1, (maxnr + 1) / 2);
    p.svd_ldt = std::max(2 * p.svd_np, 2 * rows) | 1;
    if (e->svd_accumulate_v) return 0;  // TNSU_SVD_V=accumulate (A/B comparisons) stays on the general path
    {
      const int tb = rows * p.svd_ldt + 2 * rows + 2, th = 2 * rows * 2 * rows;
      if (tb <= 2 * ldm2) {
        p.svd_tb_off = 0;
        p.svd_group_doubles = 2 * ldm2 + e->d4 + e->dcap + th;
      } else {
        p.svd_tb_off = 2 * ldm2 + e->d4 + e->dcap + th;
        p.svd_group_doubles = p.svd_tb_off + tb;
      }
    }
  } else {
    if (rows > 4) return 0;  // only thetas below 8 x 8 (or TNSU_SVD_PRECOND=0) go without the preconditioner
    p.svd_precond = 0;
    p.svd_accv = 1;
    p.svd_np = std::max(1, (maxnc + 1) / 2);
    p.svd_ldt = 0;
    p.svd_group_doubles = 2 * ldm2 + e->d4 + e->dcap + 2 * rows * 2 * p.svd_np;
  }
  p.svd_group_doubles = (p.svd_group_doubles + 1) & ~1;
  p.svd_max_sweeps = e->svd_max_sweeps;
  p.ldm = e->ldm;
  p.mrows = (maxM + 7) & ~7;
  p.ldp = maxN | 1;
  p.warp_doubles = fused_warp_doubles(p.svd_group_doubles, p.ldm, p.mrows, p.ldp);
  p.n_warps = std::min(FUSED_MAX_WARPS, width);
  std::vector<int> toff(e->n);
  long long total = 0;
  for (int t = 0; t < e->n; ++t) {
    toff[t] = (int)total;
    total += (tnsu_tensor_elems(e, t) + 1) & ~1LL;
    if (total > (1 << 20)) return 0;
  }
  p.tens_doubles = (int)total;
  long long smem = 8LL * (total + 2LL * e->m * e->dcap + e->n + 2LL * e->m + 8 + (long long)p.n_warps * p.warp_doubles);
  if (smem > SMEM_LIMIT - 1024) return 0;
  if ((long long)e->m * sizeof(EdgeDesc) <= 16384 && smem + (long long)e->m * sizeof(EdgeDesc) <= SMEM_LIMIT - 1024) {
    p.descs_in_smem = 1;
    smem += (long long)e->m * sizeof(EdgeDesc);
  }
  // The kernel is built for latency: one warp per edge, a few CTAs per SM.  It wins while the batch is at most about one
  // and a half waves of resident CTAs (measured at the config-3 shape, 4 CTAs x 148 SMs: 0.34 s against 0.42 s of the
  // launch chain at 256 points, equal at 1024, 1.9 k against 2.2 k points/s at 2048); larger batches keep the
  // general kernels, which are built for throughput.
  {
    int per_sm = 0;
    cudaError_t co = launch_run_fused_dispatch(p, rows, maxM, maxN, (int)smem, e->stream, true, &per_sm);
    if (co != cudaSuccess) return fail(e, TNSU_E_CUDA, "on-chip run kernel (occupancy query): %s", cudaGetErrorString(co));
    if (per_sm < 1) return 0;
    if (!e->force_fused && 2LL * e->batch > 3LL * per_sm * e->num_sms) return 0;
  }
  int rc = upload_site_descs(e);
  if (rc) return rc;
  CK(cudaMemcpyAsync(e->d_toff, toff.data(), sizeof(int) * e->n, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));  // toff is a local
  p.descs = e->d_descs;
  p.sites = e->d_site;
  p.level_edges = e->d_level_edges;
  p.level_off = e->d_level_off;
  p.toff = e->d_toff;
  p.n_levels = (int)e->level_edges.size();
  p.batch = e->batch;
  p.m_edges = e->m;
  p.n_tensors = e->n;
  p.dcap = e->dcap;
  p.d4 = e->d4;
  p.d = e->d;
  p.weights = e->weights;
  p.tscale = e->tscale;
  p.edge_err = e->edge_err;
  p.energy = e->energy;
  p.gates = (const double *)e->gates;
  p.gate_slot_stride = (long long)e->batch * e->m * e->d4;
  p.hams = (const double *)e->hams;
  p.rs = st;
  p.status = e->d_status;
  long long *d_prof = nullptr;
  if (getenv("TNSU_FUSED_PROFILE") != nullptr) {
    CK(cudaMalloc(&d_prof, 96 * sizeof(long long)));
    CK(cudaMemset(d_prof, 0, 96 * sizeof(long long)));
    p.prof = d_prof;
  }
  cudaError_t ce = launch_run_fused_dispatch(p, rows, maxM, maxN, (int)smem, e->stream, false, nullptr);
  if (d_prof != nullptr && ce == cudaSuccess) {
    long long h[16];
    cudaStreamSynchronize(e->stream);
    cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d_prof);
    fprintf(stderr, "[tnsu fused profile] point 0 warp 0, %lld sweeps: cycles stage %lld  lq %lld  svd %lld  recompose %lld  energy %lld  "
            "level barriers %lld | warps %d smem %lld B rows %d np %d | last SVD: sweeps %d theta+QR %d jacobi %d extract %d\n", h[7], h[0], h[1], h[2], h[3], h[4], h[5], p.n_warps, smem, rows, p.svd_np,
            ((int *)(h + 8))[7], ((int *)(h + 8))[8], ((int *)(h + 8))[9], ((int *)(h + 8))[10]);
  }
  if (ce != cudaSuccess) return fail(e, TNSU_E_CUDA, "on-chip run kernel (rows=%d smem=%lld): %s", rows, smem, cudaGetErrorString(ce));
  e->launches++;
  e->fused_runs++;
  return 1;
}

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int tnsu_abi_version(void) { return TNSU_ABI_VERSION; }

const char *tnsu_last_error(const tnsu_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int tnsu_schedule_levels(int n, int m, const int64_t *smat, int32_t *level_out) {
  if (!smat || !level_out || n <= 0 || m <= 0) return fail(nullptr, TNSU_E_ARG, "schedule_levels: bad argument");
  std::vector<int> last(n, -1);
  int nl = 0;
  for (int ek = 0; ek < m; ++ek) {
    int ends[2], cnt = 0;
    for (int t = 0; t < n; ++t)
      if (smat[(size_t)t * m + ek] > 0) {
        if (cnt < 2) ends[cnt] = t;
        ++cnt;
      }
    if (cnt != 2) return fail(nullptr, TNSU_E_ARG, "edge %d is not connected to exactly two tensors", ek);
    int lv = std::max(last[ends[0]], last[ends[1]]) + 1;
    level_out[ek] = lv;
    last[ends[0]] = last[ends[1]] = lv;
    nl = std::max(nl, lv + 1);
  }
  return nl;
}

int tnsu_destroy(tnsu_engine *e) {
  if (!e) return TNSU_OK;
  DeviceGuard guard(e->device);
  cudaStreamSynchronize(e->stream);
  for (void *p : e->tens) cudaFree(p);
  for (void *p : e->ybuf) cudaFree(p);
  if (e->small) cudaFree(e->small);
  if (e->lq_flags) cudaFree(e->lq_flags);
  if (e->lq_stats) cudaFree(e->lq_stats);
  for (int k = 0; k < 4; ++k)
    if (e->pev[k]) cudaEventDestroy(e->pev[k]);
  void *ptrs[] = {e->tnorm2, e->tticket, e->weights, e->tscale, e->edge_err, e->energy, e->gates, e->hams, e->op_buf, e->pair_val, e->cval,
                  e->rdm_buf, e->site_buf, e->d_descs, e->d_level_edges, e->d_all_edges, e->d_site, e->r_slot,
                  e->r_iter, e->r_step, e->r_active, e->r_conv, e->r_nchecks, e->r_log_slot, e->r_log_step,
                  e->r_nactive, e->r_islast, e->r_log_err, e->r_log_energy, e->dbg_sigma,
                  e->dbg_info, e->d_level_off, e->d_toff};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (e->h_nactive) cudaFreeHost(e->h_nactive);
  if (e->d_point_list) cudaFree(e->d_point_list);
  if (e->h_status) cudaFreeHost(e->h_status);
  if (e->d_status) cudaFree(e->d_status);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return TNSU_OK;
}

int tnsu_create(tnsu_engine **out, int n, int m, const int64_t *smat, int d, const int32_t *edge_dims, int d_max,
                int dtype, int batch, int device, void *stream) {
  tnsu_engine *e = nullptr;
  if (!out || !smat || !edge_dims) return fail(e, TNSU_E_ARG, "null argument");
  *out = nullptr;
  if (n <= 0 || m <= 0 || batch <= 0) return fail(e, TNSU_E_ARG, "n, m and batch must be positive");
  if (d < 1 || d > TNSU_MAX_SPIN) return fail(e, TNSU_E_UNSUPPORTED, "physical dimension %d not in 1..%d", d, TNSU_MAX_SPIN);
  if (dtype != TNSU_F64 && dtype != TNSU_C128) return fail(e, TNSU_E_ARG, "dtype must be TNSU_F64 or TNSU_C128");
  if (d_max < 1) return fail(e, TNSU_E_ARG, "d_max must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(e, TNSU_E_CUDA, "no CUDA device: the Simple-Update engine has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(e, TNSU_E_ARG, "device %d out of range (%d devices)", device, ndev);

  // structure-matrix validation (structure_matrix_constructor.py:148-178)
  std::vector<std::vector<int>> t_edges(n);
  for (int t = 0; t < n; ++t) {
    std::vector<std::pair<int, int>> ax;  // (axis, edge)
    for (int ek = 0; ek < m; ++ek) {
      int64_t a = smat[(size_t)t * m + ek];
      if (a < 0) return fail(e, TNSU_E_ARG, "negative structure-matrix entry");
      if (a > 0) ax.push_back({(int)a, ek});
    }
    std::sort(ax.begin(), ax.end());
    if ((int)ax.size() > TNSU_MAX_LEGS)
      return fail(e, TNSU_E_UNSUPPORTED, "tensor %d has %zu legs; this build holds at most %d", t, ax.size(), TNSU_MAX_LEGS);
    if (ax.empty()) return fail(e, TNSU_E_ARG, "tensor %d has no legs", t);
    for (size_t l = 0; l < ax.size(); ++l) {
      if (ax[l].first != (int)l + 1) return fail(e, TNSU_E_ARG, "row %d of the structure matrix is not a permutation of 1..z", t);
      t_edges[t].push_back(ax[l].second);
    }
  }
  std::vector<tnsu_engine::Edge> edges(m);
  for (int ek = 0; ek < m; ++ek) {
    int cnt = 0;
    for (int t = 0; t < n; ++t) {
      int64_t a = smat[(size_t)t * m + ek];
      if (a > 0) {
        if (cnt < 2) {
          edges[ek].t[cnt] = t;
          edges[ek].leg[cnt] = (int)a - 1;
        }
        ++cnt;
      }
    }
    if (cnt != 2) return fail(e, TNSU_E_ARG, "edge %d is not connected to exactly two tensors", ek);
    if (edge_dims[ek] < 1) return fail(e, TNSU_E_ARG, "edge %d has dimension %d", ek, edge_dims[ek]);
  }

  e = new tnsu_engine();
  e->device = device;
  {
    const char *env = getenv("TNSU_SVD");
    e->force_smem_svd = env && std::string(env) == "smem";
    env = getenv("TNSU_SVD_PRECOND");
    if (env && std::string(env) == "0") e->svd_precond = false;
    env = getenv("TNSU_LQ_STREAM");
    e->lq_stream = env && std::string(env) == "1";
    env = getenv("TNSU_GRAPH");
    e->use_graphs = env && std::string(env) == "1";
    env = getenv("TNSU_NCOLS");
    if (env) e->force_ncols = atoi(env);
    env = getenv("TNSU_LQ");
    e->force_unrolled_lq = env && std::string(env) == "unrolled";
    env = getenv("TNSU_SVD_V");
    e->svd_accumulate_v = env && std::string(env) == "accumulate";
    env = getenv("TNSU_LQ");
    e->force_householder = env && (std::string(env) == "householder" || std::string(env) == "unrolled");
    e->lqc_force_fallback = env && std::string(env) == "fallback";
    e->lqc_cta_kernel = env && std::string(env) == "cta";
    env = getenv("TNSU_COMPACT");
    if (env && std::string(env) == "0") e->compact_active = false;
    env = getenv("TNSU_LQ_MIN_N");
    if (env && atoi(env) > 0) e->lq_stream_min_n = atoi(env);
    env = getenv("TNSU_FUSED");
    if (env && std::string(env) == "0") e->use_fused = false;
    if (env && std::string(env) == "force") e->force_fused = true;
    env = getenv("TNSU_SVD_MAX_SWEEPS");
    if (env && atoi(env) > 0) e->svd_max_sweeps = atoi(env);
    env = getenv("TNSU_RC");
    e->force_simple_rc = env && std::string(env) == "simple";
    env = getenv("TNSU_RDM");
    e->force_simple_rdm = env && std::string(env) == "simple";
  }
  e->n = n;
  e->m = m;
  e->d = d;
  e->d4 = d * d * d * d;
  e->d_max = d_max;
  e->dtype = dtype;
  e->ssz = dtype == TNSU_C128 ? 16 : 8;
  e->batch = batch;
  e->smat.assign(smat, smat + (size_t)n * m);
  e->edges = edges;
  e->t_edges = t_edges;
  e->dims.assign(edge_dims, edge_dims + m);

  // dependency levels: an edge waits for every EARLIER edge that shares a tensor with it
  e->level.assign(m, 0);
  {
    std::vector<int> last_level_of_tensor(n, -1);
    int nl = 0;
    for (int ek = 0; ek < m; ++ek) {
      int lv = std::max(last_level_of_tensor[edges[ek].t[0]], last_level_of_tensor[edges[ek].t[1]]) + 1;
      e->level[ek] = lv;
      last_level_of_tensor[edges[ek].t[0]] = lv;
      last_level_of_tensor[edges[ek].t[1]] = lv;
      nl = std::max(nl, lv + 1);
    }
    e->level_edges.assign(nl, {});
    for (int ek = 0; ek < m; ++ek) e->level_edges[e->level[ek]].push_back(ek);
  }

  // shape trajectory -> capacities
  {
    std::vector<int> cur = e->dims;
    e->cap.assign(n, 0);
    int dcap = 1;
    auto account = [&]() {
      for (int t = 0; t < n; ++t) {
        long long el = d;
        for (int ek : t_edges[t]) el *= cur[ek];
        e->cap[t] = std::max(e->cap[t], el);
      }
      for (int ek = 0; ek < m; ++ek) dcap = std::max(dcap, cur[ek]);
    };
    account();
    for (int sweep = 0; sweep < 64; ++sweep) {
      bool changed = false;
      for (int ek = 0; ek < m; ++ek) {
        long long k[2];
        for (int s = 0; s < 2; ++s) {
          long long nprod = 1;
          for (int el : t_edges[edges[ek].t[s]])
            if (el != ek) nprod = std::min<long long>(nprod * cur[el], 1LL << 40);
          k[s] = std::min<long long>((long long)d * cur[ek], nprod);
        }
        long long dn = std::min<long long>(d_max, std::min(k[0] * d, k[1] * d));
        if (dn != cur[ek]) changed = true;
        cur[ek] = (int)dn;
        account();
        if (dcap > 64) {
          delete e;
          return fail(nullptr, TNSU_E_UNSUPPORTED, "bond dimension grows past 64");
        }
      }
      if (!changed) break;
    }
    e->dcap = dcap;
    for (int t = 0; t < n; ++t) e->cap[t] = (e->cap[t] + 1) & ~1LL;  // keep 16-byte alignment per point
  }

#define CKC(call)                                                                                \
  do {                                                                                           \
    cudaError_t _s = (call);                                                                     \
    if (_s != cudaSuccess) {                                                                     \
      fail(nullptr, TNSU_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(_s));                \
      tnsu_destroy(e);                                                                           \
      return TNSU_E_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

  DeviceGuard guard(device);
  CKC(guard.status);
  CKC(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, device));
  if (stream) {
    e->stream = (cudaStream_t)stream;
  } else {
    CKC(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->own_stream = true;
  }
  e->tens.assign(n, nullptr);
  e->ybuf.assign(n, nullptr);
  for (int t = 0; t < n; ++t) {
    CKC(cudaMalloc(&e->tens[t], (size_t)batch * e->cap[t] * e->ssz));
    CKC(cudaMemsetAsync(e->tens[t], 0, (size_t)batch * e->cap[t] * e->ssz, e->stream));
    CKC(cudaMalloc(&e->ybuf[t], (size_t)batch * e->cap[t] * e->ssz));
  }
  for (auto &lv : e->level_edges) e->max_level_size = std::max(e->max_level_size, (int)lv.size());
  for (int k = 0; k < 4; ++k) CKC(cudaEventCreate(&e->pev[k]));
  size_t B = batch;
  CKC(cudaMalloc(&e->weights, B * m * e->dcap * 8));
  CKC(cudaMemsetAsync(e->weights, 0, B * m * e->dcap * 8, e->stream));
  CKC(cudaMalloc(&e->tscale, B * n * 8));
  CKC(cudaMalloc(&e->tnorm2, B * n * 8));
  CKC(cudaMemsetAsync(e->tnorm2, 0, B * n * 8, e->stream));
  CKC(cudaMalloc(&e->tticket, B * n * 4));
  CKC(cudaMemsetAsync(e->tticket, 0, B * n * 4, e->stream));
  CKC(cudaMalloc(&e->edge_err, B * m * 8));
  CKC(cudaMemsetAsync(e->edge_err, 0, B * m * 8, e->stream));
  CKC(cudaMalloc(&e->energy, B * 8));
  CKC(cudaMalloc(&e->pair_val, B * m * 16));
  CKC(cudaMalloc(&e->cval, B * 16));
  CKC(cudaMalloc(&e->rdm_buf, B * e->d4 * 16));
  CKC(cudaMalloc(&e->site_buf, B * n * d * d * 16));
  CKC(cudaMalloc(&e->hams, B * m * e->d4 * e->ssz));
  CKC(cudaMalloc(&e->op_buf, (size_t)e->d4 * e->ssz));
  CKC(cudaMalloc(&e->d_descs, sizeof(EdgeDesc) * m));
  CKC(cudaMalloc(&e->d_site, sizeof(SiteDesc) * n));
  {
    std::vector<int> flat;
    for (auto &lv : e->level_edges) {
      e->level_off.push_back((int)flat.size());
      flat.insert(flat.end(), lv.begin(), lv.end());
    }
    CKC(cudaMalloc(&e->d_level_edges, sizeof(int) * m));
    CKC(cudaMemcpy(e->d_level_edges, flat.data(), sizeof(int) * m, cudaMemcpyHostToDevice));
    std::vector<int> loff = e->level_off;
    loff.push_back(m);
    CKC(cudaMalloc(&e->d_level_off, sizeof(int) * loff.size()));
    CKC(cudaMemcpy(e->d_level_off, loff.data(), sizeof(int) * loff.size(), cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&e->d_toff, sizeof(int) * n));
    std::vector<int> all(m);
    for (int i = 0; i < m; ++i) all[i] = i;
    CKC(cudaMalloc(&e->d_all_edges, sizeof(int) * m));
    CKC(cudaMemcpy(e->d_all_edges, all.data(), sizeof(int) * m, cudaMemcpyHostToDevice));
  }
  CKC(cudaMalloc(&e->r_slot, B * 4));
  CKC(cudaMalloc(&e->r_iter, B * 4));
  CKC(cudaMalloc(&e->r_step, B * 4));
  CKC(cudaMalloc(&e->r_active, B * 4));
  CKC(cudaMalloc(&e->r_conv, B * 4));
  CKC(cudaMalloc(&e->r_nchecks, B * 4));
  CKC(cudaMalloc(&e->r_nactive, 4));
  CKC(cudaMallocHost(&e->h_nactive, 8));  // [active points, fallback count of the streaming K1]
  CKC(cudaMalloc(&e->d_status, 2 * sizeof(int)));
  CKC(cudaMemsetAsync(e->d_status, 0, 2 * sizeof(int), e->stream));
  CKC(cudaMallocHost(&e->h_status, 2 * sizeof(int)));
  CKC(cudaMalloc(&e->dbg_sigma, 64 * 8));
  CKC(cudaMalloc(&e->dbg_info, 16 * 4));
  CKC(cudaEventCreate(&e->ev0));
  CKC(cudaEventCreate(&e->ev1));
  fill_double_kernel<<<(int)((B * n + 255) / 256), 256, 0, e->stream>>>(e->tscale, (long long)(B * n), 1.0);
  CKC(cudaGetLastError());
  CKC(cudaStreamSynchronize(e->stream));
#undef CKC
  {
    // validate eagerly: the first sweep's shapes must fit a kernel variant (later growth is checked as it happens)
    int rc = plan_sweep(e);
    if (rc) {
      g_create_error = e->err;
      tnsu_destroy(e);
      return rc;
    }
  }
  *out = e;
  return TNSU_OK;
}

int tnsu_edge_dims(const tnsu_engine *e, int32_t *dims_out) {
  if (!e || !dims_out) return TNSU_E_ARG;
  for (int i = 0; i < e->m; ++i) dims_out[i] = e->dims[i];
  return TNSU_OK;
}

int64_t tnsu_tensor_elems(const tnsu_engine *e, int t) {
  if (!e || t < 0 || t >= e->n) return -1;
  long long el = e->d;
  for (int ek : e->t_edges[t]) el *= e->dims[ek];
  return el;
}

int tnsu_num_levels(const tnsu_engine *e) { return e ? (int)e->level_edges.size() : TNSU_E_ARG; }

int tnsu_edge_levels(const tnsu_engine *e, int32_t *level_out) {
  if (!e || !level_out) return TNSU_E_ARG;
  for (int i = 0; i < e->m; ++i) level_out[i] = e->level[i];
  return TNSU_OK;
}

int tnsu_set_tensor(tnsu_engine *e, int t, const void *host, int64_t elems) {
  if (!e) return TNSU_E_ARG;
  if (!host || t < 0 || t >= e->n) return fail(e, TNSU_E_ARG, "set_tensor: bad argument");
  if (elems != tnsu_tensor_elems(e, t))
    return fail(e, TNSU_E_ARG, "set_tensor: tensor %d has %lld elements per point, got %lld", t,
                (long long)tnsu_tensor_elems(e, t), (long long)elems);
  ON_DEVICE(e);
  CK(cudaMemcpy2DAsync(e->tens[t], e->cap[t] * e->ssz, host, elems * e->ssz, elems * e->ssz, e->batch,
                       cudaMemcpyHostToDevice, e->stream));
  reset_scale_kernel<<<(e->batch + 255) / 256, 256, 0, e->stream>>>(e->tscale, e->n, e->batch, t);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  return TNSU_OK;
}

int tnsu_get_tensor(tnsu_engine *e, int t, void *host, int64_t elems) {
  if (!e) return TNSU_E_ARG;
  if (!host || t < 0 || t >= e->n) return fail(e, TNSU_E_ARG, "get_tensor: bad argument");
  if (elems != tnsu_tensor_elems(e, t))
    return fail(e, TNSU_E_ARG, "get_tensor: tensor %d has %lld elements per point, got %lld", t,
                (long long)tnsu_tensor_elems(e, t), (long long)elems);
  ON_DEVICE(e);
  dim3 grid((unsigned)std::min<long long>((elems + 255) / 256, 1024), (unsigned)std::min(e->batch, 32768));
  if (e->dtype == TNSU_C128)
    apply_scale_kernel<cplx><<<grid, 256, 0, e->stream>>>((cplx *)e->tens[t], e->cap[t], elems, e->batch, e->tscale, e->n, t);
  else
    apply_scale_kernel<double><<<grid, 256, 0, e->stream>>>((double *)e->tens[t], e->cap[t], elems, e->batch, e->tscale, e->n, t);
  CK(cudaGetLastError());
  reset_scale_kernel<<<(e->batch + 255) / 256, 256, 0, e->stream>>>(e->tscale, e->n, e->batch, t);
  CK(cudaGetLastError());
  CK(cudaMemcpy2DAsync(host, elems * e->ssz, e->tens[t], e->cap[t] * e->ssz, elems * e->ssz, e->batch,
                       cudaMemcpyDeviceToHost, e->stream));
  return poll_status(e);
}

int tnsu_set_weights(tnsu_engine *e, int edge, const double *host, int len) {
  if (!e) return TNSU_E_ARG;
  if (!host || edge < 0 || edge >= e->m) return fail(e, TNSU_E_ARG, "set_weights: bad argument");
  if (len != e->dims[edge]) return fail(e, TNSU_E_ARG, "set_weights: edge %d has dimension %d, got %d", edge, e->dims[edge], len);
  ON_DEVICE(e);
  CK(cudaMemcpy2DAsync(e->weights + (size_t)edge * e->dcap, (size_t)e->m * e->dcap * 8, host, (size_t)len * 8,
                       (size_t)len * 8, e->batch, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TNSU_OK;
}

int tnsu_get_weights(tnsu_engine *e, int edge, double *host, int len) {
  if (!e) return TNSU_E_ARG;
  if (!host || edge < 0 || edge >= e->m) return fail(e, TNSU_E_ARG, "get_weights: bad argument");
  if (len != e->dims[edge]) return fail(e, TNSU_E_ARG, "get_weights: edge %d has dimension %d, got %d", edge, e->dims[edge], len);
  ON_DEVICE(e);
  CK(cudaMemcpy2DAsync(host, (size_t)len * 8, e->weights + (size_t)edge * e->dcap, (size_t)e->m * e->dcap * 8,
                       (size_t)len * 8, e->batch, cudaMemcpyDeviceToHost, e->stream));
  return poll_status(e);
}

static int upload_c128(tnsu_engine *e, void *dst, const void *src_c128, size_t count) {
  ON_DEVICE(e);
  if (e->dtype == TNSU_C128) {
    CK(cudaMemcpyAsync(dst, src_c128, count * 16, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  } else {
    const double *s = (const double *)src_c128;
    std::vector<double> re_(count);
    for (size_t i = 0; i < count; ++i) re_[i] = s[2 * i];
    CK(cudaMemcpyAsync(dst, re_.data(), count * 8, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  return TNSU_OK;
}

int tnsu_reserve_gate_slots(tnsu_engine *e, int n_slots) {
  if (!e) return TNSU_E_ARG;
  if (n_slots < 1) return fail(e, TNSU_E_ARG, "n_slots must be >= 1");
  ON_DEVICE(e);
  CK(cudaStreamSynchronize(e->stream));
  if (e->gates) CK(cudaFree(e->gates));
  e->gates = nullptr;
  CK(cudaMalloc(&e->gates, (size_t)n_slots * e->batch * e->m * e->d4 * e->ssz));
  e->n_slots = n_slots;
  e->slot_set.assign(n_slots, 0);
  return TNSU_OK;
}

int tnsu_set_gates(tnsu_engine *e, int slot, const void *gates) {
  if (!e) return TNSU_E_ARG;
  if (!gates || slot < 0 || slot >= e->n_slots) return fail(e, TNSU_E_ARG, "set_gates: slot %d of %d", slot, e->n_slots);
  size_t count = (size_t)e->batch * e->m * e->d4;
  int rc = upload_c128(e, (char *)e->gates + (size_t)slot * count * e->ssz, gates, count);
  if (rc) return rc;
  e->slot_set[slot] = 1;
  return TNSU_OK;
}

int tnsu_set_hamiltonians(tnsu_engine *e, const void *ham) {
  if (!e) return TNSU_E_ARG;
  if (!ham) return fail(e, TNSU_E_ARG, "set_hamiltonians: null");
  int rc = upload_c128(e, e->hams, ham, (size_t)e->batch * e->m * e->d4);
  if (rc) return rc;
  e->hams_set = true;
  return TNSU_OK;
}

int tnsu_sweep(tnsu_engine *e, int slot, int n_sweeps) {
  if (!e) return TNSU_E_ARG;
  if (slot < 0 || slot >= e->n_slots || !e->slot_set[slot]) return fail(e, TNSU_E_STATE, "sweep: gates of slot %d not set", slot);
  ON_DEVICE(e);
  e->last_n = 0;
  CK(cudaEventRecord(e->ev0, e->stream));
  for (int s = 0; s < n_sweeps; ++s) {
    int rc = one_sweep(e, slot, nullptr, nullptr);
    if (rc) return rc;
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  e->timed = true;
  return TNSU_OK;
}

int tnsu_profile_sweep(tnsu_engine *e, int slot, double *ms_out) {
  if (!e) return TNSU_E_ARG;
  if (!ms_out) return fail(e, TNSU_E_ARG, "profile_sweep: null");
  if (slot < 0 || slot >= e->n_slots || !e->slot_set[slot]) return fail(e, TNSU_E_STATE, "profile_sweep: gates of slot %d not set", slot);
  ON_DEVICE(e);
  e->profile = true;
  e->prof_ms[0] = e->prof_ms[1] = e->prof_ms[2] = 0.0;
  int rc = one_sweep(e, slot, nullptr, nullptr);
  e->profile = false;
  if (rc) return rc;
  for (int k = 0; k < 3; ++k) ms_out[k] = e->prof_ms[k];
  return TNSU_OK;
}

int tnsu_last_sweep_kernel_ms(tnsu_engine *e, double *total_ms, int32_t *n_launches) {
  if (!e) return TNSU_E_ARG;
  ON_DEVICE(e);
  if (!e->timed) return fail(e, TNSU_E_STATE, "no sweep has been timed");
  CK(cudaEventSynchronize(e->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  if (total_ms) *total_ms = ms;
  if (n_launches) *n_launches = e->last_n;
  return TNSU_OK;
}

int tnsu_sweep_error(tnsu_engine *e, double *err_out) {
  if (!e) return TNSU_E_ARG;
  if (!err_out) return fail(e, TNSU_E_ARG, "sweep_error: null");
  ON_DEVICE(e);
  std::vector<double> h((size_t)e->batch * e->m);
  CK(cudaMemcpyAsync(h.data(), e->edge_err, h.size() * 8, cudaMemcpyDeviceToHost, e->stream));
  {
    int rc = poll_status(e);
    if (rc) return rc;
  }
  for (int b = 0; b < e->batch; ++b) {
    double mx = 0.0;
    bool bad = false;
    for (int ek = 0; ek < e->m; ++ek) {
      double v = h[(size_t)b * e->m + ek];
      if (v != v) bad = true;
      mx = std::max(mx, v);
    }
    err_out[b] = bad ? NAN : mx;
  }
  return TNSU_OK;
}

static int energy_launch(tnsu_engine *e, const int *iter_state, const int *active) {
  int rc;
  if (e->dtype == TNSU_C128)
    rc = launch_pair<cplx>(e, e->d_all_edges, e->m, (const cplx *)e->hams, (long long)e->m * e->d4, e->d4, e->pair_val,
                           nullptr, iter_state, active);
  else
    rc = launch_pair<double>(e, e->d_all_edges, e->m, (const double *)e->hams, (long long)e->m * e->d4, e->d4,
                             e->pair_val, nullptr, iter_state, active);
  if (rc) return rc;
  pair_sum_kernel<<<(e->batch + 127) / 128, 128, 0, e->stream>>>(e->pair_val, e->batch, e->m, e->n, e->energy, e->cval,
                                                                 iter_state, active);
  CK(cudaGetLastError());
  e->launches++;
  return TNSU_OK;
}

int tnsu_energy(tnsu_engine *e, double *energy_out) {
  if (!e) return TNSU_E_ARG;
  if (!energy_out) return fail(e, TNSU_E_ARG, "energy: null");
  if (!e->hams_set) return fail(e, TNSU_E_STATE, "energy: Hamiltonians not set");
  ON_DEVICE(e);
  int rc = ensure_static_descs(e);
  if (rc) return rc;
  rc = energy_launch(e, nullptr, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(energy_out, e->energy, (size_t)e->batch * 8, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TNSU_OK;  // NaN/Inf energies are returned as they are, like the reference does
}

int tnsu_pair_rdm(tnsu_engine *e, int edge, void *rdm) {
  if (!e) return TNSU_E_ARG;
  if (!rdm || edge < 0 || edge >= e->m) return fail(e, TNSU_E_ARG, "pair_rdm: bad argument");
  ON_DEVICE(e);
  int rc = ensure_static_descs(e);
  if (rc) return rc;
  if (e->dtype == TNSU_C128)
    rc = launch_pair<cplx>(e, e->d_all_edges + edge, 1, nullptr, 0, 0, nullptr, e->rdm_buf, nullptr, nullptr);
  else
    rc = launch_pair<double>(e, e->d_all_edges + edge, 1, nullptr, 0, 0, nullptr, e->rdm_buf, nullptr, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(rdm, e->rdm_buf, (size_t)e->batch * e->d4 * 16, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TNSU_OK;
}

static int site_rdms(tnsu_engine *e, std::vector<cplx> &h) {
  ON_DEVICE(e);
  int rc = upload_site_descs(e);
  if (rc) return rc;
  if (e->dcap > 64) return fail(e, TNSU_E_UNSUPPORTED, "site RDM: bond dimension above 64");
  if (e->dtype == TNSU_C128)
    site_rdm_kernel<cplx><<<e->batch * e->n, RDM_THREADS, 0, e->stream>>>(e->d_site, e->n, e->d, e->batch, e->m, e->dcap,
                                                                           e->weights, e->site_buf);
  else
    site_rdm_kernel<double><<<e->batch * e->n, RDM_THREADS, 0, e->stream>>>(e->d_site, e->n, e->d, e->batch, e->m, e->dcap,
                                                                             e->weights, e->site_buf);
  CK(cudaGetLastError());
  e->launches++;
  h.resize((size_t)e->batch * e->n * e->d * e->d);
  CK(cudaMemcpyAsync(h.data(), e->site_buf, h.size() * 16, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TNSU_OK;
}

int tnsu_site_rdm(tnsu_engine *e, int t, void *rdm) {
  if (!e) return TNSU_E_ARG;
  if (!rdm || t < 0 || t >= e->n) return fail(e, TNSU_E_ARG, "site_rdm: bad argument");
  std::vector<cplx> h;
  int rc = site_rdms(e, h);
  if (rc) return rc;
  const int dd = e->d * e->d;
  cplx *out = (cplx *)rdm;
  for (int b = 0; b < e->batch; ++b)
    for (int i = 0; i < dd; ++i) out[(size_t)b * dd + i] = h[((size_t)b * e->n + t) * dd + i];
  return TNSU_OK;
}

int tnsu_expectation_per_site(tnsu_engine *e, const void *op, void *out) {
  if (!e) return TNSU_E_ARG;
  if (!op || !out) return fail(e, TNSU_E_ARG, "expectation_per_site: null");
  std::vector<cplx> h;
  int rc = site_rdms(e, h);
  if (rc) return rc;
  const int d = e->d, dd = d * d;
  const cplx *o = (const cplx *)op;
  cplx *res = (cplx *)out;
  for (int b = 0; b < e->batch; ++b) {
    cplx tot = mkc(0, 0);
    for (int t = 0; t < e->n; ++t) {
      const cplx *r = &h[((size_t)b * e->n + t) * dd];
      for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) tot += r[i * d + j] * o[j * d + i];  // trace(rho @ O)
    }
    res[b] = tot * (1.0 / e->n);
  }
  return TNSU_OK;
}

int tnsu_pair_expectation_per_site(tnsu_engine *e, const void *op, void *out) {
  if (!e) return TNSU_E_ARG;
  if (!op || !out) return fail(e, TNSU_E_ARG, "pair_expectation_per_site: null");
  ON_DEVICE(e);
  int rc = ensure_static_descs(e);
  if (rc) return rc;
  rc = upload_c128(e, e->op_buf, op, (size_t)e->d4);
  if (rc) return rc;
  if (e->dtype == TNSU_C128)
    rc = launch_pair<cplx>(e, e->d_all_edges, e->m, (const cplx *)e->op_buf, 0, 0, e->pair_val, nullptr, nullptr, nullptr);
  else
    rc = launch_pair<double>(e, e->d_all_edges, e->m, (const double *)e->op_buf, 0, 0, e->pair_val, nullptr, nullptr, nullptr);
  if (rc) return rc;
  pair_sum_kernel<<<(e->batch + 127) / 128, 128, 0, e->stream>>>(e->pair_val, e->batch, e->m, e->n, nullptr, e->cval,
                                                                 nullptr, nullptr);
  CK(cudaGetLastError());
  e->launches++;
  CK(cudaMemcpyAsync(out, e->cval, (size_t)e->batch * 16, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return TNSU_OK;
}

int tnsu_run(tnsu_engine *e, int n_dts, const int32_t *is_last_value, int max_iterations, double tol, int log_cap,
             int32_t *n_checks, double *log_error, double *log_energy, int32_t *log_slot, int32_t *log_step,
             int32_t *converged, int32_t *total_sweeps, int32_t *final_slot) {
  if (!e) return TNSU_E_ARG;
  if (n_dts < 0 || n_dts > e->n_slots) return fail(e, TNSU_E_ARG, "run: %d time steps but %d gate slots", n_dts, e->n_slots);
  for (int s = 0; s < n_dts; ++s)
    if (!e->slot_set[s]) return fail(e, TNSU_E_STATE, "run: gates of slot %d not set", s);
  if (!e->hams_set) return fail(e, TNSU_E_STATE, "run: Hamiltonians not set");
  if (!is_last_value && n_dts > 0) return fail(e, TNSU_E_ARG, "run: null is_last_value");
  if (log_cap < 0) return fail(e, TNSU_E_ARG, "run: negative log_cap");
  ON_DEVICE(e);
  const size_t B = e->batch;
  // (re)allocate logs
  if (log_cap > e->r_log_cap || !e->r_log_err) {
    if (e->r_log_err) cudaFree(e->r_log_err);
    if (e->r_log_energy) cudaFree(e->r_log_energy);
    if (e->r_log_slot) cudaFree(e->r_log_slot);
    if (e->r_log_step) cudaFree(e->r_log_step);
    size_t cap = std::max(log_cap, 1);
    CK(cudaMalloc(&e->r_log_err, B * cap * 8));
    CK(cudaMalloc(&e->r_log_energy, B * cap * 8));
    CK(cudaMalloc(&e->r_log_slot, B * cap * 4));
    CK(cudaMalloc(&e->r_log_step, B * cap * 4));
    e->r_log_cap = (int)cap;
  }
  if (e->r_islast) cudaFree(e->r_islast);
  e->r_islast = nullptr;
  CK(cudaMalloc(&e->r_islast, std::max(n_dts, 1) * 4));
  if (n_dts > 0) CK(cudaMemcpy(e->r_islast, is_last_value, n_dts * 4, cudaMemcpyHostToDevice));
  const int gb = (int)((B + 255) / 256);
  fill_int_kernel<<<gb, 256, 0, e->stream>>>(e->r_slot, (int)B, 0);
  fill_int_kernel<<<gb, 256, 0, e->stream>>>(e->r_iter, (int)B, 0);
  fill_int_kernel<<<gb, 256, 0, e->stream>>>(e->r_step, (int)B, 0);
  fill_int_kernel<<<gb, 256, 0, e->stream>>>(e->r_active, (int)B, (n_dts > 0 && max_iterations > 0) ? 1 : 0);
  fill_int_kernel<<<gb, 256, 0, e->stream>>>(e->r_conv, (int)B, 0);
  fill_int_kernel<<<gb, 256, 0, e->stream>>>(e->r_nchecks, (int)B, 0);
  CK(cudaGetLastError());

  RunState st{};
  st.batch = e->batch;
  st.m_edges = e->m;
  st.n_dts = n_dts;
  st.max_iter = max_iterations;
  st.log_cap = e->r_log_cap;
  st.tol = tol;
  st.slot = e->r_slot;
  st.iter = e->r_iter;
  st.step = e->r_step;
  st.active = e->r_active;
  st.conv = e->r_conv;
  st.nchecks = e->r_nchecks;
  st.log_slot = e->r_log_slot;
  st.log_step = e->r_log_step;
  st.nactive = e->r_nactive;
  st.is_last = e->r_islast;
  st.log_err = e->r_log_err;
  st.log_energy = e->r_log_energy;
  st.edge_err = e->edge_err;
  st.energy = e->energy;

  if (e->d_point_list == nullptr) CK(cudaMalloc(&e->d_point_list, B * sizeof(int)));
  e->n_listed = -1;
  const long long max_global = (long long)n_dts * max_iterations;
  const int poll = 4;
  bool any = (n_dts > 0 && max_iterations > 0);
  // One iteration of the loop = one sweep (3 launches per dependency level) + energy (2) + the state machine (2): with
  // a few hundred points every launch is latency bound, so once the shapes are steady the iteration is captured
  // can be captured into a CUDA graph and replayed (TNSU_GRAPH=1).  Measured on the 256-point sweep: 0.58 s with the
  // graph, 0.54 s eager — the chain is bound by kernel latency on the device, not by launch overhead — so eager
  // launches stay the default.
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int64_t graph_launches = 0;
  e->lqc_suspend_until = -1;
  e->lqc_attempted = 0;
  if (e->lq_stats) CK(cudaMemsetAsync(e->lq_stats, 0, sizeof(int), e->stream));
  auto iteration = [&](long long it) -> int {
    e->run_sweep_index = it;
    int rc = one_sweep(e, 0, e->r_slot, e->r_active);
    if (rc) return rc;
    if (it >= 2) {
      // energy at check iterations (self-gated per point on r_iter); no point checks before its 3rd sweep
      rc = ensure_static_descs(e);
      if (rc) return rc;
      rc = energy_launch(e, e->r_iter, e->r_active);
      if (rc) return rc;
    }
    CK(cudaMemsetAsync(e->r_nactive, 0, 4, e->stream));
    run_state_kernel<<<gb, 256, 0, e->stream>>>(st);
    CK(cudaGetLastError());
    e->launches++;
    return TNSU_OK;
  };
  int rc_loop = TNSU_OK;
  bool fused_tried = false;
  for (long long it = 0; it < max_global && any; ++it) {
    // small networks with steady bond dimensions: the rest of the run goes to ONE persistent kernel that keeps each
    // network in shared memory (run_fused.cuh); the sweeps before (bond growth) ran through the general kernels above
    if (!e->plan_valid) {
      rc_loop = plan_sweep(e);
      if (rc_loop != TNSU_OK) break;
    }
    if (gexec == nullptr && fused_tried == false && e->plan_steady) {
      fused_tried = true;  // eligibility cannot change once the bond dimensions are steady
      const int fr = try_run_fused(e, st);
      if (fr < 0) {
        rc_loop = fr;
        break;
      }
      if (fr == 1) {
        rc_loop = poll_status(e);
        break;
      }
    }
    if (e->use_graphs && gexec == nullptr && it >= 3 && e->plan_valid && e->plan_steady) {
      const int64_t l0 = e->launches;
      if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        rc_loop = iteration(it);
        cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
        graph_launches = e->launches - l0;
        e->launches = l0;
        if (rc_loop != TNSU_OK) break;
        if (ce != cudaSuccess || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
          gexec = nullptr;
          cudaGetLastError();
          e->use_graphs = false;  // fall back to eager launches
        }
      } else {
        cudaGetLastError();
        e->use_graphs = false;
      }
    }
    if (gexec != nullptr) {
      cudaError_t ce = cudaGraphLaunch(gexec, e->stream);
      if (ce != cudaSuccess) {
        rc_loop = fail(e, TNSU_E_CUDA, "graph launch: %s", cudaGetErrorString(ce));
        break;
      }
      e->launches += graph_launches;
    } else {
      rc_loop = iteration(it);
      if (rc_loop != TNSU_OK) break;
    }
    if ((it + 1) % poll == 0 || it + 1 == max_global) {
      cudaError_t ce = cudaMemcpyAsync(e->h_nactive, e->r_nactive, 4, cudaMemcpyDeviceToHost, e->stream);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
      if (ce != cudaSuccess) {
        rc_loop = fail(e, TNSU_E_CUDA, "run loop poll: %s", cudaGetErrorString(ce));
        break;
      }
      any = (*e->h_nactive > 0);
      // points finished since the last poll: the launches from here on cover the remaining ones only (a replayed graph
      // has its grid sizes baked in and keeps the mask)
      if (any && e->compact_active && gexec == nullptr && !e->use_graphs && *e->h_nactive < (e->n_listed >= 0 ? e->n_listed : (int)B)) {
        compact_active_kernel<<<1, 1024, 0, e->stream>>>(e->r_active, (int)B, e->d_point_list);
        ce = cudaGetLastError();
        if (ce != cudaSuccess) {
          rc_loop = fail(e, TNSU_E_CUDA, "run loop compaction: %s", cudaGetErrorString(ce));
          break;
        }
        e->n_listed = *e->h_nactive;
        e->launches++;
      }
      // the streaming K1/K3 is only worth its first pass while it keeps most of the bond matrices: when more than half of
      // what it was given since the last poll came back flagged (converged states with small weights), the next 32 sweeps
      // go straight to the Householder kernels, then it is tried again
      if (e->lq_stats != nullptr && e->lqc_attempted > 0) {
        ce = cudaMemcpyAsync(e->h_nactive + 1, e->lq_stats, 4, cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaMemsetAsync(e->lq_stats, 0, sizeof(int), e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) {
          rc_loop = fail(e, TNSU_E_CUDA, "run loop poll: %s", cudaGetErrorString(ce));
          break;
        }
        if (2LL * e->h_nactive[1] > e->lqc_attempted) e->lqc_suspend_until = it + 1 + 32;
        e->lqc_flagged_total += e->h_nactive[1];
        e->lqc_attempted = 0;
      }
      rc_loop = poll_status(e);
      if (rc_loop != TNSU_OK) break;
    }
  }
  e->run_sweep_index = 0;
  e->lqc_suspend_until = -1;
  e->n_listed = -1;
  if (gexec) cudaGraphExecDestroy(gexec);
  if (graph) cudaGraphDestroy(graph);
  if (rc_loop != TNSU_OK) return rc_loop;
  CK(cudaStreamSynchronize(e->stream));
  // download per-point results
  if (n_checks) CK(cudaMemcpy(n_checks, e->r_nchecks, B * 4, cudaMemcpyDeviceToHost));
  if (converged) CK(cudaMemcpy(converged, e->r_conv, B * 4, cudaMemcpyDeviceToHost));
  if (total_sweeps) CK(cudaMemcpy(total_sweeps, e->r_step, B * 4, cudaMemcpyDeviceToHost));
  if (final_slot) CK(cudaMemcpy(final_slot, e->r_slot, B * 4, cudaMemcpyDeviceToHost));
  if (log_cap > 0) {
    size_t w = (size_t)log_cap, pitch = (size_t)e->r_log_cap;
    if (log_error) CK(cudaMemcpy2D(log_error, w * 8, e->r_log_err, pitch * 8, w * 8, B, cudaMemcpyDeviceToHost));
    if (log_energy) CK(cudaMemcpy2D(log_energy, w * 8, e->r_log_energy, pitch * 8, w * 8, B, cudaMemcpyDeviceToHost));
    if (log_slot) CK(cudaMemcpy2D(log_slot, w * 4, e->r_log_slot, pitch * 4, w * 4, B, cudaMemcpyDeviceToHost));
    if (log_step) CK(cudaMemcpy2D(log_step, w * 4, e->r_log_step, pitch * 4, w * 4, B, cudaMemcpyDeviceToHost));
  }
  return TNSU_OK;
}

void *tnsu_stream(tnsu_engine *e) { return e ? (void *)e->stream : nullptr; }

int tnsu_synchronize(tnsu_engine *e) {
  if (!e) return TNSU_E_ARG;
  ON_DEVICE(e);
  return poll_status(e);
}

int tnsu_stats(tnsu_engine *e, int64_t *out, int n_out) {
  if (!e) return TNSU_E_ARG;
  if (!out || n_out < 0) return fail(e, TNSU_E_ARG, "stats: bad argument");
  ON_DEVICE(e);
  if (e->lq_stats != nullptr) {
    int flagged = 0;
    CK(cudaMemcpyAsync(&flagged, e->lq_stats, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemsetAsync(e->lq_stats, 0, sizeof(int), e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->lqc_flagged_total += flagged;
  }
  {
    CK(cudaMemcpyAsync(e->h_status, e->d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (e->h_status[1]) {  // the non-finite flag stays for the next polling entry point
      e->svd_unconverged_total += e->h_status[1];
      CK(cudaMemsetAsync(e->d_status + 1, 0, sizeof(int), e->stream));
    }
  }
  const int64_t v[TNSU_N_STATS] = {e->launches, e->lqc_attempted_total, e->lqc_flagged_total, e->svd_unconverged_total,
                                   e->nonfinite_events, e->fused_runs};
  for (int i = 0; i < n_out && i < TNSU_N_STATS; ++i) out[i] = v[i];
  return TNSU_OK;
}

int64_t tnsu_launch_count(const tnsu_engine *e) { return e ? e->launches : -1; }

int tnsu_debug_edge(tnsu_engine *e, int slot, int edge, int point, void *l_i, void *l_j, double *sigma, int32_t *info) {
  if (!e) return TNSU_E_ARG;
  if (edge < 0 || edge >= e->m || point < 0 || point >= e->batch) return fail(e, TNSU_E_ARG, "debug_edge: bad argument");
  if (slot < 0 || slot >= e->n_slots || !e->slot_set[slot]) return fail(e, TNSU_E_STATE, "debug_edge: gates of slot %d not set", slot);
  ON_DEVICE(e);
  // A single edge out of sweep order: plan from the current dims, then only that edge's descriptor is
  // valid if no earlier edge of the plan changes a dimension it sees (true in steady state).
  if (!e->plan_valid) {
    int rc = plan_sweep(e);
    if (rc) return rc;
  }
  if (!e->plan_steady) return fail(e, TNSU_E_STATE, "debug_edge needs steady bond dimensions");
  int rc = (e->dtype == TNSU_C128) ? run_levels<cplx>(e, slot, nullptr, nullptr, edge, point)
                                   : run_levels<double>(e, slot, nullptr, nullptr, edge, point);
  if (rc) return rc;
  CK(cudaStreamSynchronize(e->stream));
  int h_info[16];
  CK(cudaMemcpy(h_info, e->dbg_info, sizeof h_info, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 16; ++i) info[i] = h_info[i];
  const int nc = std::min(h_info[4], h_info[5]);
  const int ldm = e->ldm;
  std::vector<char> tmp((size_t)ldm * ldm * e->ssz);
  auto fetch = [&](void *dst, int which, int rows, int cols) -> int {
    const char *src = (const char *)e->small + ((size_t)point * SM_COUNT + which) * ldm * ldm * e->ssz;
    CK(cudaMemcpy(tmp.data(), src, tmp.size(), cudaMemcpyDeviceToHost));
    double *o = (double *)dst;
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) {
        size_t si = (size_t)r * ldm + c, di = (size_t)r * cols + c;
        if (e->dtype == TNSU_C128) {
          o[2 * di] = ((double *)tmp.data())[2 * si];
          o[2 * di + 1] = ((double *)tmp.data())[2 * si + 1];
        } else {
          o[2 * di] = ((double *)tmp.data())[si];
          o[2 * di + 1] = 0.0;
        }
      }
    return TNSU_OK;
  };
  if ((rc = fetch(l_i, SM_L0, h_info[0], h_info[1]))) return rc;
  if ((rc = fetch(l_j, SM_L1, h_info[2], h_info[3]))) return rc;
  CK(cudaMemcpy(sigma, e->dbg_sigma, nc * 8, cudaMemcpyDeviceToHost));
  return TNSU_OK;
}

}  // extern "C"
