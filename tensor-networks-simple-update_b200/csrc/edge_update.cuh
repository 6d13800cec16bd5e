// edge_update.cuh — the per-edge imaginary-time-evolution step of Simple Update as three kernels.
//
// Replaces one iteration of the `for ek in range(m)` loop of SimpleUpdate.simple_update
// (reference src/tnsu/simple_update.py:219-296): absorb weights (tensor_network.py:242-259), permute +
// flatten (:340-369), RQ (:243-250), gate + theta (:456-485), truncated SVD (:393-409), recompose
// (:268-283), un-absorb (tensor_network.py:261-278), normalise and store (:294-296).
//
// A work item is (point b, edge e).  Its two site tensors ("sides") are independent until the gate:
//
//   K1 lq_kernel        one CTA per (item, side).  Thread t owns NCOLS whole COLUMNS of the absorbed bond
//                       matrix P (M = d*D_k rows, N = prod(other bond dims) columns) in registers:
//                         col[r] = T[s, x_k, others] * prod_{m != k} lambda_m[x_m] * tscale   (r = s*D_k + x_k)
//                       K = min(M,N) Householder steps on the ROWS of P; step k needs one block reduction
//                         g_i = sum_{c>k} P[i,c] conj(P[k,c])   over all rows i
//                       (rows i>k: reflector dot products; rows i<k: u_i^H u_k for the compact-WY factor).
//                       Out: L (M x K), S = Y1 T^H (K x K) and the reflectors Y^H (K x N, coalesced).
//   K2 svd_kernel       one small CTA per item (many resident per SM: the phase is latency bound):
//                       theta = L_i diag(lambda_k) L_j G, one-sided Jacobi SVD, truncation to D',
//                       lambda' = s / sum(s), convergence error, r~_i and r~_j.
//   K3 recompose_kernel one CTA per (item, side):  P' = r~ Q = [r~ 0] - (r~ S) Y^H  — one small GEMM per
//                       column, no reductions — times prod lambda_m^-1, stored IN PLACE (every read of T
//                       happened in K1); ||T'||^2 is block-reduced and kept as the lazy scale `tscale`
//                       that the next load applies (T / ||T||, reference :294-295).
//
// Gauge: any orthonormal row basis Q is admissible (SURVEY.md §8a "Gauge facts"); weights, RDMs and
// energies are gauge invariant and are what the parity tests compare.
#pragma once
#include <cooperative_groups.h>
#include <type_traits>

#include "common.cuh"

// compile-time loop: f(std::integral_constant<int, I>) for I in [0, N) — guarantees constant indices into
// register-resident arrays where `#pragma unroll` on a large body is only a hint
template <int I, int N, typename F>
DEV void static_for(F &&f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

#define TNSU_SMEM_LIMIT (227 * 1024)
#define SVD_THREADS 128
#define SVD_LPP 8  // lanes cooperating on one column pair

// Small per-item matrices kept in global memory between the three kernels: [item][6][ldm*ldm]
enum { SM_L0 = 0, SM_L1 = 1, SM_S0 = 2, SM_S1 = 3, SM_R0 = 4, SM_R1 = 5, SM_COUNT = 6 };

template <typename S>
struct EdgeLaunch {
  const EdgeDesc *descs;     // indexed by edge id
  const int *level_edges;    // edge ids handled by this launch
  int n_level_edges;
  const int *point_list;     // nullptr: launch position = point; else the points this launch covers (batch = its length):
                             // tnsu_run compacts the points still active so that finished ones cost nothing
  int batch, m_edges, dcap, d4, n_tensors;
  double *weights;           // [batch][m][dcap]
  double *tscale;            // [batch][n_tensors] lazy 1/||T||
  const S *gates;            // [slot][batch][m][d4]
  long long gate_slot_stride;
  const int *point_slot;     // per-point slot (run loop) or nullptr
  int fixed_slot;
  const int *active;         // per-point flag or nullptr
  double *edge_err;          // [batch][m]
  S *small;                  // [batch * n_level_edges][SM_COUNT][ldm*ldm]
  int ldm;                   // leading dimension of the small matrices
  int lda, ldv;              // leading dimensions of the SVD work arrays (column-major)
  int nr_max;                // largest theta row count (after the rows >= cols transposition) in this launch
  int nt;                    // threads per side CTA (multiple of 32)
  int lq_rolled;             // real float64: lq_rolled_kernel instead of lq_kernel
  int cluster;               // CTAs per (item, side) in K1 (thread-block cluster) and K3 (column slices); 1 = one CTA
  double *tnorm2;            // [batch][n_tensors] partial ||T'||^2 of the K3 column slices (cluster > 1)
  unsigned *tticket;         // [batch][n_tensors] arrival counter of the K3 column slices
  int lq_stream_grid;        // > 0: persistent streaming variant (TMA-staged tensors) with this many CTAs
  int lq_stage_off;          // byte offset of the staging buffer in dynamic shared memory
  int lq_stage_doubles;      // its size
  int svdw_np;               // register-resident warp SVD (svd_warp.cuh): processors (column pairs) per theta, 0 = not used
  int svdw_rows;             // rows of [A] per lane = template R of svd_warp_kernel
  int svdw_group_doubles;    // shared-memory doubles per theta in svd_warp_kernel
  int svdw_precond;          // pivoted-QR preconditioner before the Jacobi iteration
  int svdw_ldt;              // leading dimension of its transpose buffer
  const int *lq_flags;       // streaming K1/K3 (lq_chol.cuh) ran first: per (item, side) 1 = the Householder kernels must process it, 0 = skip; nullptr = process all
  int *lq_stats;             // [1] number of (item, side)s the Householder fallback had to factor (run-loop statistics) or nullptr
  int lqc_force_fallback;    // TNSU_LQ=fallback: the streaming kernel flags every (item, side) (tests of the fallback path)
  int svdw_accv;             // 1: accumulate the right rotations in registers; 0: V-free iteration, vectors recovered from theta
  int svdw_tb_off;           // V-free variant: offset (doubles, from the group's base) of the transpose buffer + pivot column
  int smem_lq, smem_svd, smem_rc;
  int svd_vfree;             // shared-memory SVD kernel: iterate A alone, recover the kept right vectors from the initial A afterwards
  int svd_precond;           // shared-memory SVD kernel (V-free only): pivoted Householder QR first, Jacobi on R^H
  int svd_max_sweeps;        // Jacobi sweep limit (60 for the warp kernel, 40 for the shared-memory one; TNSU_SVD_MAX_SWEEPS overrides: tests)
  int *status;               // [2] device failure counters: [0] != 0 once a theta had no finite positive singular-value sum
                             // (NaN / Inf in the state), [1] Jacobi iterations that hit their sweep limit
  // debug dump (tnsu_debug_edge)
  double *dbg_sigma;
  int *dbg_info;
  int dbg_point;
};

// ---------------------------------------------------------------------------------------------
// warp reduce-scatter of NV per-lane values: afterwards lane l holds the warp total of value (l % NV)
// ---------------------------------------------------------------------------------------------
template <typename S, int NV>
DEV S warp_reduce_scatter(S (&v)[NV], int lane) {
  static_assert(NV == 1 || NV == 2 || NV == 4 || NV == 8 || NV == 16 || NV == 32, "NV must be a power of two <= 32");
  static_for<0, 5>([&](auto st) {
    constexpr int half = (NV >> 1) >> decltype(st)::value;  // NV/2, NV/4, ... 1
    if constexpr (half >= 1) {
      const bool upper = (lane & half) != 0;  // lanes with bit `half` set keep the upper half of the live values
#pragma unroll
      for (int j = 0; j < half; ++j) {
        S keep = upper ? v[j + half] : v[j];
        S send = upper ? v[j] : v[j + half];
        v[j] = keep + shfl_xor(send, half);
      }
    }
  });
#pragma unroll
  for (int m = NV; m < 32; m *= 2) v[0] = v[0] + shfl_xor(v[0], m);
  return v[0];
}

// column c of the bond matrix -> element offsets of its (s=0, x_k=0) entry before / after the update and the
// product of the environment weights on the other legs
struct ColInfo {
  long long in_off, out_off;
  double w, inv_w;
};
// n / d and n % d for n < 2^22, d <= 1024 through a float reciprocal (one MUFU + a few integer instructions instead of
// the ~25-instruction integer division): the truncated estimate is at most one off in either direction
DEV void divmod_small(unsigned n, unsigned d, unsigned &q, unsigned &r) {
  q = (unsigned)(__uint2float_rz(n) * __frcp_rn(__uint2float_rz(d)));
  int rr = (int)n - (int)(q * d);
  if (rr < 0) {
    --q;
    rr += (int)d;
  } else if (rr >= (int)d) {
    ++q;
    rr -= (int)d;
  }
  r = (unsigned)rr;
}

DEV ColInfo decode_column(const SideDesc &sd, int c, const double *wleg, int dcap) {
  ColInfo ci;
  ci.in_off = 0;
  ci.out_off = 0;
  ci.w = 1.0;
  ci.inv_w = 1.0;
  unsigned rem = (unsigned)c;
  for (int l = sd.nlegs - 1; l >= 0; --l) {
    if (l == sd.bond_leg) continue;
    const int dim = sd.dims[l];
    unsigned x;
    divmod_small(rem, (unsigned)dim, rem, x);
    ci.in_off += (long long)x * sd.in_stride[l];
    ci.out_off += (long long)x * sd.out_stride[l];
    const double wl_x = wleg[l * dcap + x];
    ci.w *= wl_x;
    ci.inv_w *= 1.0 / wl_x;  // the reference multiplies by np.power(w, -1) leg by leg (tensor_network.py:274)
  }
  return ci;
}

template <typename S, int MMAX, int NCOLS>
constexpr int edge_max_threads() {
  return (MMAX * NCOLS * (int)sizeof(S) / 8 > 48) ? 256 : 512;
}

// =============================================================================================
// K1: absorb + Householder LQ + compact-WY factors
// =============================================================================================
// CLUSTER = true (wide bond matrices, complex128: the real path has its own cluster variant in lq_rolled_kernel): the
// columns of one (item, side) are split over the p.cluster CTAs of a thread-block cluster; each step's partial Gram row
// goes to the CTA's shared memory, one cluster.sync(), every CTA sums the peers' rows through distributed shared memory
// and reads the pivot column from the first CTA.
template <typename S, int MMAX, int NCOLS, bool CLUSTER = false>
__global__ void __launch_bounds__(edge_max_threads<S, MMAX, NCOLS>(), 1) lq_kernel(const EdgeLaunch<S> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CL = CLUSTER ? p.cluster : 1;
  const int rank = CLUSTER ? (int)(blockIdx.x % CL) : 0;
  const int item = (int)(blockIdx.x / CL) >> 1;
  const int side = (int)(blockIdx.x / CL) & 1;
  const int bi = item / p.n_level_edges;  // position in the launch; the point it stands for comes from the list of active points
  const int ek = p.level_edges[item - bi * p.n_level_edges];
  const int b = p.point_list != nullptr ? p.point_list[bi] : bi;
  if (p.active != nullptr && p.active[b] == 0) return;
  // behind the streaming CholeskyQR2 kernel (lq_chol.cuh): only the (item, side)s it flagged are factored here
  if (p.lq_flags != nullptr) {
    if (p.lq_flags[blockIdx.x / CL] == 0) return;
    if (p.lq_stats != nullptr && threadIdx.x == 0 && rank == 0) atomicAdd(p.lq_stats, 1);
  }

  const EdgeDesc &ed = p.descs[ek];
  const SideDesc &sd = ed.s[side];
  const int nt = blockDim.x;
  const int lt = threadIdx.x;
  const int lane = lt & 31;
  const int wl = lt >> 5;
  const int nwarps = nt >> 5;
  const int M = sd.M, N = sd.N, K = sd.K;
  const int Dk = ed.Dk;
  const int ldm = p.ldm;
  constexpr int CH = MMAX < 8 ? MMAX : 8;

  // shared memory: wleg | rowin | red | colk | matL | matZ | matT | matY | tau
  double *wleg = (double *)smem_raw;                                  // [MAX_LEGS][dcap]
  long long *rowin = (long long *)(wleg + TNSU_MAX_LEGS * p.dcap);    // [MMAX]
  S *red = (S *)(rowin + MMAX);                                       // [2][TNSU_MAX_WARPS][MMAX]
  S *colk = red + 2 * TNSU_MAX_WARPS * MMAX;                          // [2][MMAX]
  S *matL = colk + 2 * MMAX;                                          // L (M x K)
  S *matZ = matL + ldm * ldm;                                         // Z (K x K), later S = Y1 T^H
  S *matT = matZ + ldm * ldm;                                         // T (K x K)
  S *matY = matT + ldm * ldm;                                         // first K columns of Y^H
  double *tau_s = (double *)(matY + ldm * ldm);                       // [MMAX]
  S *cpart = (S *)(smem_raw + p.lq_stage_off);                        // [2][MMAX] this CTA's partial Gram row (CLUSTER)
  const int ncl = (N + CL - 1) / CL;                                   // columns per CTA of the cluster
  const int c0 = rank * ncl;
  const int c_end = min(N, c0 + ncl);

  const double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  for (int idx = lt; idx < sd.nlegs * p.dcap; idx += nt) {
    int l = idx / p.dcap, x = idx - l * p.dcap;
    wleg[idx] = (x < sd.dims[l]) ? wpoint[(size_t)sd.leg_edge[l] * p.dcap + x] : 1.0;
  }
  for (int r = lt; r < MMAX; r += nt) {
    int s = r / Dk, x = r - s * Dk;
    rowin[r] = (r < M) ? (long long)s * sd.in_stride_phys + (long long)x * sd.in_stride[sd.bond_leg] : 0;
  }
  for (int idx = lt; idx < ldm * ldm; idx += nt) matL[idx] = from_real<S>(0.0);
  __syncthreads();

  // ---- load my columns, absorbing the environment weights ----
  const S *tbase = (const S *)sd.base + (size_t)b * sd.point_stride;
  const double tsc = p.tscale[(size_t)b * p.n_tensors + sd.tensor];
  S col[NCOLS][MMAX];
#pragma unroll
  for (int j = 0; j < NCOLS; ++j) {
    const int c = c0 + lt + j * nt;
    const bool valid = c < c_end;
    const ColInfo ci = decode_column(sd, valid ? c : 0, wleg, p.dcap);
    const double wc = ci.w * tsc;
#pragma unroll
    for (int r = 0; r < MMAX; ++r) {
      S v = from_real<S>(0.0);
      if (valid && r < M) v = tbase[ci.in_off + rowin[r]] * wc;
      col[j][r] = v;
    }
  }

  // ---- Householder LQ on the rows of P; reflector k stays in col[][k] ----
  static_for<0, MMAX>([&](auto kk_c) {
    constexpr int kk = decltype(kk_c)::value;
    if (kk < K) {
      constexpr int buf = kk & 1;
#pragma unroll
      for (int j = 0; j < NCOLS; ++j) {
        if (c0 + lt + j * nt == kk) {
#pragma unroll
          for (int i = 0; i < MMAX; ++i) colk[buf * MMAX + i] = col[j][i];
        }
      }
      // partial Gram row over my columns, rows in chunks of CH to bound live registers
#pragma unroll
      for (int cb = 0; cb < MMAX; cb += CH) {
        if (cb < M) {
          S acc[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) acc[i] = from_real<S>(0.0);
#pragma unroll
          for (int j = 0; j < NCOLS; ++j) {
            if (c0 + lt + j * nt > kk) {
              const S pk = cj(col[j][kk]);
#pragma unroll
              for (int i = 0; i < CH; ++i) fma_acc(acc[i], col[j][cb + i], pk);
            }
          }
          S tot = warp_reduce_scatter<S, CH>(acc, lane);
          if (lane < CH) red[(buf * TNSU_MAX_WARPS + wl) * MMAX + cb + lane] = tot;
        }
      }
      __syncthreads();

      // every warp finalises redundantly: lane l owns row l
      S g = from_real<S>(0.0);
      S mycol = from_real<S>(0.0);
      if (lane < M) {
        for (int w = 0; w < nwarps; ++w) g = g + red[(buf * TNSU_MAX_WARPS + w) * MMAX + lane];
      }
      const S *colk_src = colk;
      if constexpr (CLUSTER) {
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        if (wl == 0 && lane < MMAX) cpart[buf * MMAX + lane] = g;
        cluster.sync();
        S tot = from_real<S>(0.0);
        if (lane < M)
          for (int r = 0; r < CL; ++r) tot = tot + cluster.map_shared_rank(cpart, r)[buf * MMAX + lane];
        g = tot;
        colk_src = cluster.map_shared_rank(colk, 0);
      }
      if (lane < M) mycol = colk_src[buf * MMAX + lane];
      const S alpha = colk_src[buf * MMAX + kk];
      const double gk = re(shfl_idx(g, kk));
      const double a2 = abs2(alpha);
      const double norm2 = gk + a2;
      S w_l = from_real<S>(0.0), t_l = g, lnew = mycol, head = from_real<S>(0.0), lkk = from_real<S>(0.0);
      double tau = 0.0;
      if (norm2 > 0.0) {
        const double norm = sqrt(norm2);
        const double absa = sqrt(a2);
        const S phase_c = (absa > 0.0) ? cj(alpha) * (1.0 / absa) : from_real<S>(1.0);
        const S uk = cj(alpha) + phase_c * norm;  // head of the reflector u_k
        tau = 1.0 / (norm * (norm + absa));
        t_l = g + mycol * uk;
        w_l = t_l * tau;
        lnew = mycol - w_l * cj(uk);
        head = cj(uk);
        lkk = -(cj(phase_c) * norm);
      }
      if (wl == 0 && lane < M) {
        if (lane > kk)
          matL[lane * ldm + kk] = lnew;
        else if (lane == kk) {
          matL[kk * ldm + kk] = lkk;
          tau_s[kk] = tau;
        } else
          matZ[lane * ldm + kk] = t_l;
      }
      // rank-1 update of rows i > kk of my columns; fix up row kk (the stored reflector)
#pragma unroll
      for (int i = kk + 1; i < MMAX; ++i) {
        const S wi = shfl_idx(w_l, i);
#pragma unroll
        for (int j = 0; j < NCOLS; ++j) {
          if (c0 + lt + j * nt > kk) fms_acc(col[j][i], wi, col[j][kk]);
        }
      }
#pragma unroll
      for (int j = 0; j < NCOLS; ++j) {
        const int c = c0 + lt + j * nt;
        if (c == kk)
          col[j][kk] = head;
        else if (c < kk)
          col[j][kk] = from_real<S>(0.0);
      }
    }
  });

  // reflectors to global (coalesced rows of Y^H) and their first K columns to shared memory
  S *ybase = (S *)sd.ybase + (size_t)b * sd.point_stride;
#pragma unroll
  for (int j = 0; j < NCOLS; ++j) {
    const int c = c0 + lt + j * nt;
    if (c < c_end) {
#pragma unroll
      for (int k = 0; k < MMAX; ++k)
        if (k < K) ybase[(size_t)k * N + c] = col[j][k];
    }
    if (c < K) {
#pragma unroll
      for (int k = 0; k < MMAX; ++k)
        if (k < K) matY[k * ldm + c] = col[j][k];
    }
  }
  __syncthreads();
  if constexpr (CLUSTER) {
    // no CTA may leave while a peer can still read its shared memory; only the first CTA writes L and S
    cooperative_groups::this_cluster().sync();
    if (rank != 0) return;
  }

  // compact-WY factor: T[:k,k] = -tau_k T[:k,:k] Z[:k,k], T[k,k] = tau_k   (lane r holds row r)
  if (wl == 0) {
    S trow[MMAX];
#pragma unroll
    for (int k = 0; k < MMAX; ++k) trow[k] = from_real<S>(0.0);
    static_for<0, MMAX>([&](auto kk_c) {
      constexpr int kk = decltype(kk_c)::value;
      if (kk < K) {
        S a = from_real<S>(0.0);
#pragma unroll
        for (int j = 0; j < kk; ++j) fma_acc(a, trow[j], matZ[j * ldm + kk]);
        const double tk = tau_s[kk];
        trow[kk] = (lane < kk) ? -(a * tk) : (lane == kk ? from_real<S>(tk) : from_real<S>(0.0));
      }
    });
    if (lane < K) {
#pragma unroll
      for (int k = 0; k < MMAX; ++k)
        if (k < K) matT[lane * ldm + k] = trow[k];
    }
  }
  __syncthreads();
  // S[c][k'] = (Y1 T^H)[c][k'] = conj( sum_k Yh[k][c] T[k'][k] ) and L, both to the item's global scratch
  S *small = p.small + (size_t)item * SM_COUNT * ldm * ldm;
  S *gS = small + (SM_S0 + side) * ldm * ldm;
  S *gL = small + (SM_L0 + side) * ldm * ldm;
  for (int idx = lt; idx < K * K; idx += nt) {
    const int c = idx / K, kp = idx - c * K;
    S a = from_real<S>(0.0);
    for (int k = kp; k <= c; ++k) fma_acc(a, matY[k * ldm + c], matT[kp * ldm + k]);
    gS[c * ldm + kp] = cj(a);
  }
  for (int idx = lt; idx < M * K; idx += nt) {
    const int r = idx / K, q = idx - r * K;
    gL[r * ldm + q] = matL[r * ldm + q];
  }
}

// =============================================================================================
// K1 (real float64), rolled form: the same Householder LQ with the step loop NOT unrolled over the rows.
// After every block of U steps the register rows rotate by U, so inside a block the pivot row is a
// compile-time register (row u), the rows still to be reduced follow it and the finished reflectors queue
// up at the end of the register column — their inner products with the new reflector (compact-WY T factor)
// fall out of the same reduction.  ~1 500 SASS instructions instead of ~10 000: the unrolled kernel spent
// 18 % of its issue slots waiting for instruction fetch (profiles/r01a).  The per-step reduction of the MMAX
// Gram values is a shared-memory transpose inside the warp (MMAX stores + MMAX loads per lane) followed by one
// block barrier for the cross-warp sum.
// =============================================================================================
#define LQR_U 4       // steps per rotation block
#define LQR_TPLD 33   // leading dimension of the per-warp transpose buffer

// mbarrier / bulk-copy (TMA) helpers for the streaming variant
DEV unsigned smem_u32(const void *ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
DEV void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DEV void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEV void bulk_g2s(void *dst_smem, const void *src_global, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_global), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
DEV void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// STREAM = true: persistent CTAs (grid = resident CTAs) loop over the (item, side) work list; the site tensor of the
// NEXT work item is copied global -> shared by one TMA bulk copy (cp.async.bulk + mbarrier) while the Householder
// steps of the current one run, so the 64 KiB load latency that opened every CTA of the one-shot kernel
// (19 % of its stall samples, profiles/r01e) is off the critical path.
// optional per-phase cycle counters of the Householder step (compile with -DLQR_PROFILE; read through tnsu_debug_edge,
// info[11..15] = dots+transpose-store | warp sum | barrier | cross-warp sum+scalar | update+store, summed over the steps)
#ifdef LQR_PROFILE
#define LQR_CLK(name) const long long name = clock64()
#define LQR_ACC() do { prof[0] += c1 - c0; prof[1] += c2 - c1; prof[2] += c3 - c2; prof[3] += c4 - c3; prof[4] += c5 - c4; } while (0)
#define LQR_DUMP() do { if (p.dbg_info != nullptr && b == p.dbg_point && side == 0 && lt == 0) for (int i_ = 0; i_ < 5; ++i_) p.dbg_info[11 + i_] = (int)prof[i_]; } while (0)
#else
#define LQR_CLK(name)
#define LQR_ACC()
#define LQR_DUMP()
#endif

template <int MMAX, int NCOLS, bool STREAM, bool CLUSTER>
__global__ void __launch_bounds__(edge_max_threads<double, MMAX, NCOLS>(), 1) lq_rolled_kernel(const EdgeLaunch<double> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nt = blockDim.x;
  const int lt = threadIdx.x;
  const int lane = lt & 31;
  const int wl = lt >> 5;
  const int nwarps = nt >> 5;
  const int ldm = p.ldm;
  constexpr int PARTS = 32 / MMAX;  // lanes sharing one register row in the transpose reduction
  const int q = lane & (MMAX - 1);  // the register row this lane finalises
  const int part = lane / MMAX;
  (void)PARTS;

  // shared memory: wleg | rowin | red | colk | matL | matZ | matT | matY | tau | tp[nwarps] | (stream: stage | mbarrier)
  double *wleg = (double *)smem_raw;                                  // [MAX_LEGS][dcap]
  int *rowin = (int *)(wleg + TNSU_MAX_LEGS * p.dcap);                // [MMAX]
  double *red = (double *)(rowin + MMAX + (MMAX & 1));                // [2][TNSU_MAX_WARPS][MMAX]
  double *colk = red + 2 * TNSU_MAX_WARPS * MMAX;                     // [2][MMAX]
  double *matL = colk + 2 * MMAX;                                     // L (M x K)
  double *matZ = matL + ldm * ldm;                                    // Z (K x K), later S = Y1 T^H
  double *matT = matZ + ldm * ldm;                                    // T (K x K)
  double *matY = matT + ldm * ldm;                                    // first K columns of Y^H
  double *tau_s = matY + ldm * ldm;                                   // [MMAX]
  double *tp = tau_s + MMAX + (size_t)wl * MMAX * LQR_TPLD;           // [MMAX][LQR_TPLD] per warp
  double *stage = (double *)(smem_raw + p.lq_stage_off);              // staged site tensor (STREAM)
  unsigned long long *bar = (unsigned long long *)(stage + p.lq_stage_doubles);

  // thread-block cluster (large bond matrices): CL CTAs share one (item, side); CTA `rank` owns columns
  // [rank*ncl, rank*ncl + ncl) and the per-step Gram row is summed across the cluster through distributed shared memory
  const int CL = CLUSTER ? p.cluster : 1;  // compile-time 1 in the single-CTA variant: all cluster code folds away
  const int rank = CLUSTER ? (int)(blockIdx.x % CL) : 0;
  double *cpart = (double *)(smem_raw + p.lq_stage_off);              // [2][MMAX] this CTA's partial Gram row (CL > 1)
  const int n_work = 2 * p.batch * p.n_level_edges;
  unsigned parity = 0;
  // issue the bulk copy of work item w2's tensor (thread 0 only); returns false when that item is skipped
  auto issue = [&](int w2) {
    const int item2 = w2 >> 1, side2 = w2 & 1;
    const int bi2 = item2 / p.n_level_edges;
    const int b2 = p.point_list != nullptr ? p.point_list[bi2] : bi2;
    if (p.active != nullptr && p.active[b2] == 0) return;
    const EdgeDesc &ed2 = p.descs[p.level_edges[item2 - bi2 * p.n_level_edges]];
    const SideDesc &sd2 = ed2.s[side2];
    const unsigned bytes = (unsigned)((sd2.in_stride_phys * ed2.d * 8 + 15) & ~15LL);
    mbar_expect_tx(bar, bytes);
    bulk_g2s(stage, (const double *)sd2.base + (size_t)b2 * sd2.point_stride, bytes, bar);
  };
  if constexpr (STREAM) {
    if (lt == 0) {
      mbar_init(bar, 1);
      fence_proxy_async();
    }
    __syncthreads();
    if (lt == 0 && (int)blockIdx.x < n_work) issue(blockIdx.x);
  }

  // one work item per CTA, except the persistent variants: STREAM, and the fallback behind the streaming
  // CholeskyQR2 kernel (lq_flags), where a small grid scans the flags and factors only the flagged (item, side)s
  const bool looped = STREAM || (!CLUSTER && p.lq_flags != nullptr);
  int n_fallback = 0;  // flagged work items this CTA factored (lq_flags mode)
  for (int w = (int)(blockIdx.x / CL); w < n_work; w += (looped ? (int)gridDim.x : n_work)) {
  if (!CLUSTER && p.lq_flags != nullptr && p.lq_flags[w] == 0) continue;  // taken by the streaming kernel: one load per scanned item
  const int item = w >> 1;
  const int side = w & 1;
  const int bi = item / p.n_level_edges;  // position in the launch; the point it stands for comes from the list of active points
  const int ek = p.level_edges[item - bi * p.n_level_edges];
  const int b = p.point_list != nullptr ? p.point_list[bi] : bi;
  if (p.active != nullptr && p.active[b] == 0) {
    if constexpr (STREAM) {
      // nothing was staged for this item; stage the next one now (the buffer is free)
      if (lt == 0 && w + (int)gridDim.x < n_work) issue(w + gridDim.x);
      continue;
    } else {
      if (looped) continue;
      return;
    }
  }

  if (p.lq_flags != nullptr && p.lq_flags[w] == 0) continue;  // factored by the streaming CholeskyQR2 kernel already
  ++n_fallback;

  const EdgeDesc &ed = p.descs[ek];
  const SideDesc &sd = ed.s[side];
  const int M = sd.M, N = sd.N, K = sd.K;
  const int Dk = ed.Dk;
  const int ncl = (N + CL - 1) / CL;           // columns per CTA of the cluster
  const int c0 = rank * ncl;                   // my first column
  const int c_end = min(N, c0 + ncl);

  const double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  for (int idx = lt; idx < sd.nlegs * p.dcap; idx += nt) {
    int l = idx / p.dcap, x = idx - l * p.dcap;
    wleg[idx] = (x < sd.dims[l]) ? wpoint[(size_t)sd.leg_edge[l] * p.dcap + x] : 1.0;
  }
  for (int r = lt; r < MMAX; r += nt) {
    int s = r / Dk, x = r - s * Dk;
    rowin[r] = (r < M) ? (int)(s * sd.in_stride_phys + x * sd.in_stride[sd.bond_leg]) : 0;
  }
  for (int idx = lt; idx < ldm * ldm; idx += nt) matL[idx] = 0.0;

  // ---- load my columns (raw: the loads are in flight while the weights arrive in shared memory) ----
  if constexpr (STREAM) {
    mbar_wait(bar, parity);
    parity ^= 1u;
  }
  const double *tbase = STREAM ? (const double *)stage : (const double *)sd.base + (size_t)b * sd.point_stride;
  const double tsc = p.tscale[(size_t)b * p.n_tensors + sd.tensor];
  double col[NCOLS][MMAX];
  bool cvalid[NCOLS];
#pragma unroll
  for (int j = 0; j < NCOLS; ++j) {
    const int c = c0 + lt + j * nt;
    cvalid[j] = c < c_end;
    // element offset of (s = 0, x_k = 0) of column c: mixed-radix decode over the other legs
    long long off = 0;
    unsigned rem = cvalid[j] ? (unsigned)c : 0u;
    for (int l = sd.nlegs - 1; l >= 0; --l) {
      if (l == sd.bond_leg) continue;
      unsigned x;
      divmod_small(rem, (unsigned)sd.dims[l], rem, x);
      off += (long long)x * sd.in_stride[l];
    }
    const int bstride = (int)sd.in_stride[sd.bond_leg];
    const long long pstride = sd.in_stride_phys;
    // row r = (s, x_k): walk (s, x_k) instead of dividing by Dk per row
    const double *rp = tbase + off;
    int xk = 0;
#pragma unroll
    for (int r = 0; r < MMAX; ++r) {
      col[j][r] = (cvalid[j] && r < M) ? rp[(long long)xk * bstride] : 0.0;
      if (++xk == Dk) {
        xk = 0;
        rp += pstride;
      }
    }
  }
  __syncthreads();
  if constexpr (STREAM) {
    // every thread has copied its columns out of the staging buffer: refill it for this CTA's next work item
    if (lt == 0 && w + (int)gridDim.x < n_work) {
      fence_proxy_async();
      issue(w + gridDim.x);
    }
  }
  // ---- absorb the environment weights and the lazy normalisation ----
#pragma unroll
  for (int j = 0; j < NCOLS; ++j) {
    const ColInfo ci = decode_column(sd, cvalid[j] ? c0 + lt + j * nt : 0, wleg, p.dcap);
    const double wc = ci.w * tsc;
#pragma unroll
    for (int r = 0; r < MMAX; ++r) col[j][r] *= wc;
  }
  double *ybase = (double *)sd.ybase + (size_t)b * sd.point_stride;

  // ---- Householder steps, LQR_U per rotation block ----
  long long prof[5] = {0, 0, 0, 0, 0};
  (void)prof;
#pragma unroll 1
  for (int kk0 = 0; kk0 < K; kk0 += LQR_U) {
    static_for<0, LQR_U>([&](auto u_c) {
      constexpr int u = decltype(u_c)::value;
      const int kk = kk0 + u;
      if (kk < K) {
        const int buf = kk & 1;
        LQR_CLK(c0);
        if (rank == 0 && lt == kk) {  // column kk is column j = 0 of thread kk of the first CTA (kk < 32 <= nt)
#pragma unroll
          for (int i = 0; i < MMAX; ++i) colk[buf * MMAX + i] = col[0][i];
        }
        // partial Gram row of the pivot (register row u) with every register row, over my columns c > kk
        double pk[NCOLS];
#pragma unroll
        for (int j = 0; j < NCOLS; ++j) pk[j] = (rank > 0 || j > 0 || lt > kk) ? col[j][u] : 0.0;
#pragma unroll
        for (int i = 0; i < MMAX; ++i) {
          double a = col[0][i] * pk[0];
#pragma unroll
          for (int j = 1; j < NCOLS; ++j) a = fma(col[j][i], pk[j], a);
          tp[i * LQR_TPLD + lane] = a;
        }
        __syncwarp();
        LQR_CLK(c1);
        // lane (q, part) sums MMAX of the 32 lane partials of register row q, then the parts combine
        double g;
        {
          const double *src = tp + q * LQR_TPLD + part * MMAX;
          double s0 = src[0], s1 = src[1], s2 = src[2], s3 = src[3];
#pragma unroll
          for (int i = 4; i < MMAX; i += 4) {
            s0 += src[i];
            s1 += src[i + 1];
            s2 += src[i + 2];
            s3 += src[i + 3];
          }
          g = (s0 + s1) + (s2 + s3);
#pragma unroll
          for (int m = MMAX; m < 32; m *= 2) g += __shfl_xor_sync(0xffffffffu, g, m);
        }
        LQR_CLK(c2);
        if (lane < MMAX) red[(buf * TNSU_MAX_WARPS + wl) * MMAX + q] = g;
        __syncthreads();
        LQR_CLK(c3);
        {
          const double *rr = red + (size_t)buf * TNSU_MAX_WARPS * MMAX + q;
          double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
          int w = 0;
          for (; w + 4 <= nwarps; w += 4) {
            g0 += rr[w * MMAX];
            g1 += rr[(w + 1) * MMAX];
            g2 += rr[(w + 2) * MMAX];
            g3 += rr[(w + 3) * MMAX];
          }
          for (; w < nwarps; ++w) g0 += rr[w * MMAX];
          g = (g0 + g1) + (g2 + g3);
        }
        const double *colk_src = colk;
        if constexpr (CLUSTER) {
          // sum the CTA partials across the cluster; column kk lives in the first CTA
          namespace cg = cooperative_groups;
          cg::cluster_group cluster = cg::this_cluster();
          if (wl == 0 && lane < MMAX) cpart[buf * MMAX + q] = g;
          cluster.sync();
          double tot = 0.0;
          for (int r = 0; r < CL; ++r) tot += cluster.map_shared_rank(cpart, r)[buf * MMAX + q];
          g = tot;
          colk_src = cluster.map_shared_rank(colk, 0);
        }
        const double mycol = colk_src[buf * MMAX + q];
        const double alpha = colk_src[buf * MMAX + u];
        const double gk = __shfl_sync(0xffffffffu, g, u);
        const double norm2 = fma(alpha, alpha, gk);
        // roles of register row q: q < u reflector kk0+q; q == u pivot; u < q < MMAX-kk0 matrix row kk0+q;
        // q >= MMAX-kk0 reflector q-(MMAX-kk0)
        const int n_mat = MMAX - kk0;
        const bool is_row = q > u && q < n_mat;
        double w_l = 0.0, t_l = g, lnew = mycol, head = 0.0, lkk = 0.0, tau = 0.0;
        if (norm2 > 0.0) {
          const double norm = sqrt(norm2);
          const double absa = fabs(alpha);
          const double phase = (alpha < 0.0) ? -1.0 : 1.0;
          const double uk = fma(phase, norm, alpha);  // head of the reflector u_k
          tau = 1.0 / (norm * (norm + absa));
          t_l = fma(mycol, uk, g);
          w_l = t_l * tau;
          lnew = fma(-w_l, uk, mycol);
          head = uk;
          lkk = -(phase * norm);
        }
        if (wl == 0 && lane < MMAX) {
          if (q == u) {
            matL[kk * ldm + kk] = lkk;
            tau_s[kk] = tau;
          } else if (is_row) {
            if (kk0 + q < M) matL[(kk0 + q) * ldm + kk] = lnew;
          } else {
            const int ri = (q < u) ? kk0 + q : q - n_mat;
            matZ[ri * ldm + kk] = t_l;
          }
        }
        if (!is_row) w_l = 0.0;
        LQR_CLK(c4);
        // rank-1 update of the matrix rows still to be reduced (other register rows get w = 0)
#pragma unroll
        for (int i = u + 1; i < MMAX; ++i) {
          const double wi = __shfl_sync(0xffffffffu, w_l, i);
#pragma unroll
          for (int j = 0; j < NCOLS; ++j) col[j][i] = fma(-wi, pk[j], col[j][i]);
        }
        // the pivot row becomes the stored reflector: row kk of Y^H goes to global (coalesced) and to matY
#pragma unroll
        for (int j = 0; j < NCOLS; ++j) {
          const int c = c0 + lt + j * nt;
          const double y = (c == kk) ? head : (c < kk ? 0.0 : col[j][u]);
          col[j][u] = y;
          if (cvalid[j]) ybase[(size_t)kk * N + c] = y;
          if (rank == 0 && j == 0 && lt < K) matY[kk * ldm + lt] = y;
        }
        LQR_CLK(c5);
        LQR_ACC();
      }
    });
    // rotate the register rows by LQR_U: the block's reflectors move behind the matrix rows
#pragma unroll
    for (int j = 0; j < NCOLS; ++j) {
      double t[LQR_U];
#pragma unroll
      for (int i = 0; i < LQR_U; ++i) t[i] = col[j][i];
#pragma unroll
      for (int i = 0; i + LQR_U < MMAX; ++i) col[j][i] = col[j][i + LQR_U];
#pragma unroll
      for (int i = 0; i < LQR_U; ++i) col[j][MMAX - LQR_U + i] = t[i];
    }
  }
  __syncthreads();
  LQR_DUMP();
  if constexpr (CLUSTER) {
    // no CTA may leave while a peer can still read its shared memory; only the first CTA writes L and S
    cooperative_groups::this_cluster().sync();
    if (rank != 0) return;
  }

  // compact-WY factor: T[:k,k] = -tau_k T[:k,:k] Z[:k,k], T[k,k] = tau_k   (lane r holds row r)
  if (wl == 0) {
    double trow[MMAX];
#pragma unroll
    for (int k = 0; k < MMAX; ++k) trow[k] = 0.0;
    static_for<0, MMAX>([&](auto kk_c) {
      constexpr int kk = decltype(kk_c)::value;
      if (kk < K) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // four partial sums: the chain is latency bound
#pragma unroll
        for (int j = 0; j < kk; ++j) {
          const double z = matZ[j * ldm + kk];
          if ((j & 3) == 0) a0 = fma(trow[j], z, a0);
          if ((j & 3) == 1) a1 = fma(trow[j], z, a1);
          if ((j & 3) == 2) a2 = fma(trow[j], z, a2);
          if ((j & 3) == 3) a3 = fma(trow[j], z, a3);
        }
        const double a = (a0 + a1) + (a2 + a3);
        const double tk = tau_s[kk];
        trow[kk] = (lane < kk) ? -(a * tk) : (lane == kk ? tk : 0.0);
      }
    });
    if (lane < K) {
#pragma unroll
      for (int k = 0; k < MMAX; ++k)
        if (k < K) matT[lane * ldm + k] = trow[k];
    }
  }
  __syncthreads();
  // S[c][k'] = (Y1 T^H)[c][k'] = sum_k Yh[k][c] T[k'][k]  and L, both to the item's global scratch
  double *small = p.small + (size_t)item * SM_COUNT * ldm * ldm;
  double *gS = small + (SM_S0 + side) * ldm * ldm;
  double *gL = small + (SM_L0 + side) * ldm * ldm;
  for (int idx = lt; idx < K * K; idx += nt) {
    const int c = idx / K, kp = idx - c * K;
    double a = 0.0;
    for (int k = kp; k <= c; ++k) a = fma(matY[k * ldm + c], matT[kp * ldm + k], a);
    gS[c * ldm + kp] = a;
  }
  for (int idx = lt; idx < M * K; idx += nt) {
    const int r = idx / K, qq = idx - r * K;
    gL[r * ldm + qq] = matL[r * ldm + qq];
  }
  if (looped) __syncthreads();  // the shared matrices are reused by the next work item
  }  // work loop
  if (p.lq_flags != nullptr && p.lq_stats != nullptr && lt == 0 && n_fallback > 0) atomicAdd(p.lq_stats, n_fallback);
}

// =============================================================================================
// K2: theta + one-sided Jacobi SVD + truncation
// =============================================================================================
DEV double rot_sign(double x) { return x < 0.0 ? -1.0 : 1.0; }

// rotation (c, s e^{-i phi}, s e^{+i phi}) that orthogonalises two columns with Gram entries al, be, ga = a_p^H a_q
DEV void jacobi_rotation(double al, double be, double ga, double &c, double &sp, double &sq) {
  const double delta = be - al;
  const double inv_r = rsqrt(delta * delta + 4.0 * ga * ga);
  const double c2 = 0.5 + 0.5 * fabs(delta) * inv_r;
  const double inv_c = rsqrt(c2);
  c = c2 * inv_c;
  const double s = rot_sign(delta) * ga * inv_r * inv_c;  // carries the sign of gamma: real e^{i phi} = +-1
  sp = s;
  sq = s;
}
DEV void jacobi_rotation(double al, double be, cplx ga, double &c, cplx &sp, cplx &sq) {
  const double g2 = abs2(ga);
  const double delta = be - al;
  const double inv_r = rsqrt(delta * delta + 4.0 * g2);
  const double c2 = 0.5 + 0.5 * fabs(delta) * inv_r;
  const double inv_c = rsqrt(c2);
  c = c2 * inv_c;
  const double f = rot_sign(delta) * inv_r * inv_c;  // s e^{i phi} = f * gamma
  sq = ga * f;
  sp = cj(ga) * f;
}

// A is nr x nc column-major (nr >= nc), V nc x nc.  RPL = rows of A per lane, CPL = rows of V per lane.
template <typename S, int RPL, int CPL>
__device__ int jacobi_svd(S *A, int lda, int nr, int nc, S *V, int ldv, volatile int *flag, int tid, int nthr, int max_sweeps,
                          bool accv) {
  // accv = false: the right rotations are not accumulated (V is not touched): the caller recovers the vectors it needs
  if (accv)
    for (int idx = tid; idx < nc * nc; idx += nthr) {
      int col = idx / nc, row = idx - col * nc;
      V[col * ldv + row] = from_real<S>(row == col ? 1.0 : 0.0);
    }
  const int lane8 = tid & (SVD_LPP - 1);
  const int group = tid / SVD_LPP;
  const int ngroups = nthr / SVD_LPP;
  const unsigned gmask = 0xFFu << (8 * ((tid & 31) >> 3));
  const int np = (nc + 1) & ~1;  // even number of players; index >= nc is a bye
  const int npairs = np / 2;
  const int nm1 = np - 1;
  const double tol2 = 2.25e-30;  // (1.5e-15)^2
  int sweeps = 0;
  __syncthreads();
  // flag[0]: a rotation happened in this sweep; flag[2], flag[3]: the largest binary exponents of cos^2 of a pair met
  // and of |sin|^2 of a rotation applied.  As in the warp kernel (svd_warp.cuh) the iteration also ends after a sweep
  // whose (max cosine) x (max |sin|) is below 7.9e-16: the cosines such a sweep leaves behind are <= 2 (n - 1) times
  // that, 1e-13 at n = 64, and the confirming sweep would only establish it.
  int *iflag = const_cast<int *>(flag);
  for (; sweeps < max_sweeps; ++sweeps) {
    if (tid == 0) {
      flag[0] = 0;
      flag[2] = -100000;
      flag[3] = -100000;
    }
    __syncthreads();
    for (int r = 0; r < nm1; ++r) {
      for (int pp = group; pp < npairs; pp += ngroups) {
        // round-robin tournament: player np-1 is fixed, the others rotate
        int p = r + pp, q = r - pp + nm1;
        if (p >= nm1) p -= nm1;
        if (q >= nm1) q -= nm1;
        if (pp == 0) {
          p = nm1;
          q = r;
        }
        if (p > q) {
          int t = p;
          p = q;
          q = t;
        }
        if (q < nc) {
          S *ap = A + (size_t)p * lda;
          S *aq = A + (size_t)q * lda;
          S *vp = V + (size_t)p * ldv;
          S *vq = V + (size_t)q * ldv;
          S x[RPL], y[RPL], vx[CPL], vy[CPL];
#pragma unroll
          for (int u = 0; u < RPL; ++u) {
            const int i = lane8 + u * SVD_LPP;
            x[u] = (i < nr) ? ap[i] : from_real<S>(0.0);
            y[u] = (i < nr) ? aq[i] : from_real<S>(0.0);
          }
#pragma unroll
          for (int u = 0; u < CPL; ++u) {
            const int i = lane8 + u * SVD_LPP;
            vx[u] = (accv && i < nc) ? vp[i] : from_real<S>(0.0);
            vy[u] = (accv && i < nc) ? vq[i] : from_real<S>(0.0);
          }
          double al = 0.0, be = 0.0;
          S ga = from_real<S>(0.0);
#pragma unroll
          for (int u = 0; u < RPL; ++u) {
            al += abs2(x[u]);
            be += abs2(y[u]);
            fma_acc(ga, cj(x[u]), y[u]);
          }
#pragma unroll
          for (int m = SVD_LPP / 2; m >= 1; m >>= 1) {
            al += __shfl_xor_sync(gmask, al, m);
            be += __shfl_xor_sync(gmask, be, m);
            ga = ga + shfl_xor(ga, m, gmask);
          }
          const double g2 = abs2(ga);
          if (g2 > tol2 * al * be && g2 > 0.0) {
            double c;
            S sp, sq;
            jacobi_rotation(al, be, ga, c, sp, sq);
#pragma unroll
            for (int u = 0; u < RPL; ++u) {
              const int i = lane8 + u * SVD_LPP;
              if (i < nr) {
                ap[i] = x[u] * c - sp * y[u];
                aq[i] = sq * x[u] + y[u] * c;
              }
            }
            if (accv) {
#pragma unroll
              for (int u = 0; u < CPL; ++u) {
                const int i = lane8 + u * SVD_LPP;
                if (i < nc) {
                  vp[i] = vx[u] * c - sp * vy[u];
                  vq[i] = sq * vx[u] + vy[u] * c;
                }
              }
            }
            if (lane8 == 0) {
              flag[0] = 1;
              atomicMax(iflag + 2, ((__double2hiint(g2) >> 20) & 0x7ff) - ((__double2hiint(al * be) >> 20) & 0x7ff));
              atomicMax(iflag + 3, ((__double2hiint(abs2(sq)) >> 20) & 0x7ff) - 1023);
            }
          }
        }
      }
      __syncthreads();
    }
    if (flag[0] == 0 || flag[2] + flag[3] <= -103) return sweeps + 1;
    __syncthreads();  // everyone has read the flag before the next sweep resets it
  }
  return 99;  // sweep limit reached without a rotation-free sweep
}

// resident CTAs per SM the register allocator should aim for: the Jacobi phase is latency bound, so occupancy
// is what buys throughput; 4*RPL staged scalars plus ~44 bookkeeping registers per thread
template <typename S, int RPL>
constexpr int svd_min_blocks() {
  const int need = 4 * RPL * (int)sizeof(S) / 4 + 44;
  const int b = 65536 / (SVD_THREADS * need);
  return b < 1 ? 1 : (b > 12 ? 12 : b);
}

template <typename S, int RPL>
__global__ void __launch_bounds__(SVD_THREADS, svd_min_blocks<S, RPL>()) svd_kernel(const EdgeLaunch<S> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int item = blockIdx.x;
  const int bi = item / p.n_level_edges;  // position in the launch; the point it stands for comes from the list of active points
  const int ek = p.level_edges[item - bi * p.n_level_edges];
  const int b = p.point_list != nullptr ? p.point_list[bi] : bi;
  if (p.active != nullptr && p.active[b] == 0) return;
  const EdgeDesc &ed = p.descs[ek];
  const SideDesc &si = ed.s[0];
  const SideDesc &sj = ed.s[1];
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const int d = ed.d, Dk = ed.Dk, Dnew = ed.Dnew;
  const int Ki = si.K, Kj = sj.K;
  const int ni = ed.ni, nj = ed.nj;
  const int ldm = p.ldm;
  const bool transposed = ni < nj;  // Jacobi wants rows >= cols
  const int nr = transposed ? nj : ni;
  const int nc = transposed ? ni : nj;
  long long clk0 = clock64();

  // shared memory: Li | Lj | gate | svdA | svdV | sig | lam_old | order | flag
  S *Li = (S *)smem_raw;
  S *Lj = Li + ldm * ldm;
  S *gate_s = Lj + ldm * ldm;
  S *svdA = gate_s + p.d4;
  // V (nc x nc) overlays L_i | L_j, which are dead once theta is built, whenever it fits there (a third less shared
  // memory per theta: 4 instead of 3 resident CTAs per SM at 32 x 32 complex128)
  // V-free iteration (p.svd_vfree): a copy A0 of the initial A sits where V would, nothing is accumulated, and the Dnew
  // right vectors that survive the truncation are recovered at the end as w_x = A0^H u_x / ||A0^H u_x|| (+ two passes of
  // modified Gram-Schmidt) into Wk [dcap][ldv].  The accumulator is 2/7 of the iteration's shared-memory traffic and
  // flops, and this kernel runs at 73 % of the shared-memory bandwidth (profiles/r02).
  const bool vfree = p.svd_vfree != 0;
  const bool v_on_l = !vfree && (size_t)p.ldv * p.ldv <= (size_t)2 * ldm * ldm;
  S *svdV = v_on_l ? Li : svdA + (size_t)p.lda * p.ldv;  // nc <= ldv columns
  S *A0 = svdA + (size_t)p.lda * p.ldv;                  // V-free: [nc][lda] like A
  S *Wk = A0 + (size_t)p.lda * p.ldv;                    // V-free: [Dnew][lda]
  double *sig = (double *)(svdA + (size_t)p.lda * p.ldv +
                           (vfree ? (size_t)p.lda * p.ldv + (size_t)p.dcap * p.lda : (v_on_l ? 0 : (size_t)p.ldv * p.ldv)));
  double *lam_old = sig + TNSU_MAX_SVD;
  int *order = (int *)(lam_old + TNSU_MAX_SVD);
  int *flag = order + TNSU_MAX_SVD;

  S *small = p.small + (size_t)item * SM_COUNT * ldm * ldm;
  double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  {
    const S *gLi = small + SM_L0 * ldm * ldm, *gLj = small + SM_L1 * ldm * ldm;
    for (int idx = tid; idx < si.M * ldm; idx += nthr) Li[idx] = gLi[idx];
    for (int idx = tid; idx < sj.M * ldm; idx += nthr) Lj[idx] = gLj[idx];
    const int slot = p.point_slot ? p.point_slot[b] : p.fixed_slot;
    const S *gsrc = p.gates + (size_t)slot * p.gate_slot_stride + ((size_t)b * p.m_edges + ek) * p.d4;
    for (int idx = tid; idx < p.d4; idx += nthr) gate_s[idx] = gsrc[idx];
    for (int x = tid; x < Dk; x += nthr) lam_old[x] = wpoint[(size_t)ek * p.dcap + x];
  }
  __syncthreads();

  // theta[(qi,i'),(j',qj)] = sum_{i,j,e} L_i[(i,e),qi] lam_e L_j[(j,e),qj] G[i,j,i',j']
  for (int idx = tid; idx < Ki * Kj; idx += nthr) {
    const int qi = idx / Kj, qj = idx - qi * Kj;
    S x[TNSU_MAX_SPIN][TNSU_MAX_SPIN];
#pragma unroll
    for (int i = 0; i < TNSU_MAX_SPIN; ++i)
#pragma unroll
      for (int j = 0; j < TNSU_MAX_SPIN; ++j) {
        S a = from_real<S>(0.0);
        if (i < d && j < d)
          for (int e = 0; e < Dk; ++e) fma_acc(a, Li[(i * Dk + e) * ldm + qi] * lam_old[e], Lj[(j * Dk + e) * ldm + qj]);
        x[i][j] = a;
      }
#pragma unroll
    for (int ip = 0; ip < TNSU_MAX_SPIN; ++ip)
#pragma unroll
      for (int jp = 0; jp < TNSU_MAX_SPIN; ++jp) {
        if (ip < d && jp < d) {
          S a = from_real<S>(0.0);
#pragma unroll
          for (int i = 0; i < TNSU_MAX_SPIN; ++i)
#pragma unroll
            for (int j = 0; j < TNSU_MAX_SPIN; ++j)
              if (i < d && j < d) fma_acc(a, x[i][j], gate_s[((i * d + j) * d + ip) * d + jp]);
          const int row = qi * d + ip, cl = jp * Kj + qj;
          if (transposed)
            svdA[(size_t)row * p.lda + cl] = cj(a);  // A = theta^H
          else
            svdA[(size_t)cl * p.lda + row] = a;
        }
      }
  }
  long long clk1 = clock64();
  __syncthreads();  // theta is complete; every read of L_i / L_j is done before V is written on top of them
  if (vfree) {
    for (int idx = tid; idx < nc * p.lda; idx += nthr) A0[idx] = svdA[idx];
  }
  // Preconditioner (Drmac-Veselic, as in svd_warp.cuh): Householder QR with column pivoting A = Q R~ (columns stay in
  // place, R~ = R P^T), then the Jacobi iteration runs on X = R~^H (nc x nc), whose columns are graded: 13 -> ~8 sweeps.
  // With the V-free recovery Q is never needed: X W = Y gives the nc-side singular vectors of A as the normalised
  // columns of Y (row index = A's own column index), and the nr-side ones are recovered as A0 y / ||A0 y|| below.
  const bool precond = vfree && p.svd_precond != 0;
  int nr_it = nr;  // rows of the matrix the Jacobi iteration sees
  if (precond) {
    double *cn = sig;        // trailing squared column norms (sig / order are free until the iteration is over)
    int *step_of = order;    // the step at which a column became the pivot, -1 while it waits
    S *wv = Wk;              // the current reflector (Wk is free until the recovery)
    __syncthreads();
    for (int c = tid; c < nc; c += nthr) {
      double a = 0.0;
      for (int r = 0; r < nr; ++r) a += abs2(svdA[(size_t)c * p.lda + r]);
      cn[c] = a;
      step_of[c] = -1;
    }
    const int lane = tid & 31, wid = tid >> 5;
    __syncthreads();
    for (int k = 0; k < nc; ++k) {
      if (wid == 0) {
        // pivot = the waiting column with the largest trailing norm (ties: lowest index)
        double best = -1.0;
        int bi = 0;
        for (int c = lane; c < nc; c += 32)
          if (step_of[c] < 0 && cn[c] > best) {
            best = cn[c];
            bi = c;
          }
#pragma unroll
        for (int mm = 16; mm >= 1; mm >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, mm);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, mm);
          if (ob > best || (ob == best && oi < bi)) {
            best = ob;
            bi = oi;
          }
        }
        S *ap = svdA + (size_t)bi * p.lda;
        double n2 = 0.0;
        for (int r = k + lane; r < nr; r += 32) n2 += abs2(ap[r]);
#pragma unroll
        for (int mm = 16; mm >= 1; mm >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, mm);
        const S alpha = ap[k];
        __syncwarp();
        double tau = 0.0;
        if (n2 > 0.0) {
          const double norm = sqrt(n2);
          const double absa = sqrt(abs2(alpha));
          const S phase = (absa > 0.0) ? alpha * (1.0 / absa) : from_real<S>(1.0);
          tau = 1.0 / (norm * (norm + absa));
          for (int r = k + 1 + lane; r < nr; r += 32) {
            wv[r] = ap[r];
            ap[r] = from_real<S>(0.0);
          }
          if (lane == 0) {
            wv[k] = alpha + phase * norm;
            ap[k] = -(phase * norm);
          }
        }
        if (lane == 0) {
          step_of[bi] = k;
          flag[1] = bi;
          ((volatile double *)cn)[bi] = tau;  // the pivot's norm slot is free now: carries tau to the other warps
        }
      }
      __syncthreads();
      {
        const int piv = flag[1];
        const double tau = cn[piv];
        if (tau != 0.0) {
          const int lane8 = tid & (SVD_LPP - 1), group = tid / SVD_LPP, ngroups = nthr / SVD_LPP;
          const unsigned gmask = 0xFFu << (8 * ((tid & 31) >> 3));
          for (int c0 = 0; c0 < nc; c0 += ngroups) {  // every lane of a warp takes part in the shuffles of each trip
            const int c = c0 + group;
            const bool live = c < nc && step_of[c < nc ? c : 0] < 0;
            S *ac = svdA + (size_t)(c < nc ? c : 0) * p.lda;
            S dot = from_real<S>(0.0);
            if (live)
              for (int r = k + lane8; r < nr; r += SVD_LPP) fma_acc(dot, cj(wv[r]), ac[r]);
#pragma unroll
            for (int mm = SVD_LPP / 2; mm >= 1; mm >>= 1) dot = dot + shfl_xor(dot, mm, gmask);
            const S f = dot * tau;
            double trail = 0.0;
            if (live)
              for (int r = k + lane8; r < nr; r += SVD_LPP) {
                S v = ac[r];
                fms_acc(v, f, wv[r]);
                ac[r] = v;
                if (r > k) trail += abs2(v);
              }
#pragma unroll
            for (int mm = SVD_LPP / 2; mm >= 1; mm >>= 1) trail += __shfl_xor_sync(gmask, trail, mm);
            if (live && lane8 == 0) cn[c] = trail;
          }
        }
      }
      __syncthreads();
    }
    // X = R~^H in place: conjugate transpose of the top nc x nc block; entries below a column's pivot row are exact zeros
    for (int idx = tid; idx < nc * nc; idx += nthr) {
      const int i = idx / nc, c = idx - i * nc;
      if (i <= c) {
        const S a = (i <= step_of[c]) ? svdA[(size_t)c * p.lda + i] : from_real<S>(0.0);  // R~[i][c]
        const S bq = (c <= step_of[i]) ? svdA[(size_t)i * p.lda + c] : from_real<S>(0.0);  // R~[c][i]
        svdA[(size_t)i * p.lda + c] = cj(a);  // X[c][i]
        svdA[(size_t)c * p.lda + i] = cj(bq);  // X[i][c]
      }
    }
    nr_it = nc;
    __syncthreads();
  }
  const int svd_sweeps = jacobi_svd<S, RPL, RPL>(svdA, p.lda, nr_it, nc, svdV, p.ldv, flag, tid, nthr, min(p.svd_max_sweeps, 40), !vfree);
  long long clk2 = clock64();

  // singular values, descending order
  for (int j = tid; j < nc; j += nthr) {
    double s2 = 0.0;
    for (int i = 0; i < nr_it; ++i) s2 += abs2(svdA[(size_t)j * p.lda + i]);
    sig[j] = sqrt(s2);
  }
  __syncthreads();
  for (int j = tid; j < nc; j += nthr) {
    const double sj_ = sig[j];
    int pos = 0;
    for (int i = 0; i < nc; ++i) {
      const double si_ = sig[i];
      pos += (si_ > sj_ || (si_ == sj_ && i < j)) ? 1 : 0;
    }
    order[pos] = j;
  }
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int x = 0; x < Dnew; ++x) tot += sig[order[x]];
    if (!(tot > 0.0 && tot < 1e300)) atomicOr(p.status, 1);
    if (svd_sweeps == 99) atomicAdd(p.status + 1, 1);
    double err2 = 0.0;
    for (int x = 0; x < Dnew; ++x) {
      const double ln = sig[order[x]] / tot;
      wpoint[(size_t)ek * p.dcap + x] = ln;
      const double df = ln - lam_old[x];
      err2 += df * df;
    }
    p.edge_err[(size_t)b * p.m_edges + ek] = (Dnew == Dk) ? sqrt(err2) : nan("");
  }
  if (vfree) {
    // without preconditioner: w_x[c] = sum_r conj(A0[r][c]) u_x[r], u_x = (column order[x] of the final A) / sigma_x
    // (nc entries); with it the iteration delivered the nc-side vectors y_x and the nr-side ones are recovered:
    // w_x[r] = sum_c A0[r][c] y_x[c] / sigma_x (nr entries)
    const int nw = precond ? nr : nc;  // length of the recovered vectors
    for (int idx = tid; idx < Dnew * nw; idx += nthr) {
      const int xp = idx / nw, c = idx - xp * nw;
      const int oc = order[xp];
      const S *ax = svdA + (size_t)oc * p.lda;
      S acc0 = from_real<S>(0.0), acc1 = from_real<S>(0.0);
      if (precond) {
        const S *a0 = A0 + c;  // row c of A0
        int q = 0;
        for (; q + 1 < nc; q += 2) {
          fma_acc(acc0, a0[(size_t)q * p.lda], ax[q]);
          fma_acc(acc1, a0[(size_t)(q + 1) * p.lda], ax[q + 1]);
        }
        if (q < nc) fma_acc(acc0, a0[(size_t)q * p.lda], ax[q]);
      } else {
        const S *a0 = A0 + (size_t)c * p.lda;
        int r = 0;
        for (; r + 1 < nr; r += 2) {
          fma_acc(acc0, cj(a0[r]), ax[r]);
          fma_acc(acc1, cj(a0[r + 1]), ax[r + 1]);
        }
        if (r < nr) fma_acc(acc0, cj(a0[r]), ax[r]);
      }
      const double sv = sig[oc];
      Wk[(size_t)xp * p.lda + c] = (acc0 + acc1) * (sv > 0.0 ? 1.0 / sv : 0.0);
    }
    __syncthreads();
    // two passes of modified Gram-Schmidt in rank order; warp w projects the vectors j + 1 + w, j + 1 + w + nwarps, ...
    const int lane = tid & 31, wid = tid >> 5, nwarps = nthr >> 5;
    for (int pass = 0; pass < 2; ++pass) {
      for (int j = 0; j < Dnew; ++j) {
        S *wj = Wk + (size_t)j * p.lda;
        if (wid == 0) {
          double n2 = 0.0;
          for (int c = lane; c < nw; c += 32) n2 += abs2(wj[c]);
#pragma unroll
          for (int mm = 16; mm >= 1; mm >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, mm);
          if (!(n2 > 1e-280)) {
            // no direction of its own (sigma = 0, or A0^H u underflows): a fixed dense start vector, as in svd_warp.cuh
            for (int c = lane; c < nw; c += 32)
              wj[c] = from_real<S>((double)(((unsigned)(c + 1) * 2654435761u ^ (unsigned)(j + 1) * 40503u) >> 8 & 0xffffu) - 32767.5);
            __syncwarp();
            n2 = 0.0;
            for (int c = lane; c < nw; c += 32) n2 += abs2(wj[c]);
#pragma unroll
            for (int mm = 16; mm >= 1; mm >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, mm);
          }
          const double sc = rsqrt(n2);
          for (int c = lane; c < nw; c += 32) wj[c] = wj[c] * sc;
        }
        __syncthreads();
        for (int i = j + 1 + wid; i < Dnew; i += nwarps) {
          S *wi = Wk + (size_t)i * p.lda;
          S dot = from_real<S>(0.0);
          for (int c = lane; c < nw; c += 32) fma_acc(dot, cj(wj[c]), wi[c]);
#pragma unroll
          for (int mm = 16; mm >= 1; mm >>= 1) dot = dot + shfl_xor(dot, mm);
          for (int c = lane; c < nw; c += 32) fms_acc(wi[c], dot, wj[c]);
        }
        __syncthreads();
      }
    }
  }
  // r~_i[(i',x'),qi] = U[(qi,i'),x'],   r~_j[(j',x'),qj] = V^H[x',(j',qj)]
  S *rt_i = small + SM_R0 * ldm * ldm, *rt_j = small + SM_R1 * ldm * ldm;
  for (int idx = tid; idx < d * Dnew * Ki; idx += nthr) {
    const int row = idx / Ki, qi = idx - row * Ki;
    const int ip = row / Dnew, xp = row - ip * Dnew;
    const int oc = order[xp];
    const double sv = sig[oc];
    const double isv = sv > 0.0 ? 1.0 / sv : 0.0;
    const int trow = qi * d + ip;
    // the iterated matrix carries the nr-side index without the preconditioner and the nc-side index with it; the other
    // side is the recovered (V-free) or accumulated vector
    const S aval = svdA[(size_t)oc * p.lda + trow] * isv;
    const S vval = vfree ? Wk[(size_t)xp * p.lda + trow] : svdV[(size_t)oc * p.ldv + trow];
    rt_i[row * ldm + qi] = (transposed != precond) ? vval : aval;
  }
  for (int idx = tid; idx < d * Dnew * Kj; idx += nthr) {
    const int row = idx / Kj, qj = idx - row * Kj;
    const int jp = row / Dnew, xp = row - jp * Dnew;
    const int oc = order[xp];
    const double sv = sig[oc];
    const double isv = sv > 0.0 ? 1.0 / sv : 0.0;
    const int tcol = jp * Kj + qj;
    const S aval = svdA[(size_t)oc * p.lda + tcol] * isv;
    const S vval = vfree ? Wk[(size_t)xp * p.lda + tcol] : svdV[(size_t)oc * p.ldv + tcol];
    rt_j[row * ldm + qj] = (transposed != precond) ? cj(aval) : cj(vval);
  }
  if (p.dbg_info != nullptr && b == p.dbg_point) {
    for (int j = tid; j < nc; j += nthr) p.dbg_sigma[j] = sig[order[j]];
    if (tid == 0) {
      p.dbg_info[0] = si.M;
      p.dbg_info[1] = Ki;
      p.dbg_info[2] = sj.M;
      p.dbg_info[3] = Kj;
      p.dbg_info[4] = ni;
      p.dbg_info[5] = nj;
      p.dbg_info[6] = Dnew;
      p.dbg_info[7] = svd_sweeps;
      p.dbg_info[8] = (int)(clk1 - clk0);          // theta
      p.dbg_info[9] = (int)(clk2 - clk1);          // jacobi
      p.dbg_info[10] = (int)(clock64() - clk2);    // extract
    }
  }
}

// =============================================================================================
// K3: recompose, un-absorb, store in place, norm
// =============================================================================================
template <typename S, int MMAX, int NCOLS>
__global__ void __launch_bounds__(edge_max_threads<S, MMAX, NCOLS>(), 1) recompose_kernel(const EdgeLaunch<S> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int CL = p.cluster;  // column slices per (item, side)
  const int rank = (int)(blockIdx.x % CL);
  const int item = (int)(blockIdx.x / CL) >> 1;
  const int side = (int)(blockIdx.x / CL) & 1;
  const int bi = item / p.n_level_edges;  // position in the launch; the point it stands for comes from the list of active points
  const int ek = p.level_edges[item - bi * p.n_level_edges];
  const int b = p.point_list != nullptr ? p.point_list[bi] : bi;
  if (p.active != nullptr && p.active[b] == 0) return;
  if (p.lq_flags != nullptr && p.lq_flags[blockIdx.x / CL] == 0) return;  // recomposed by recompose_chol_kernel
  const EdgeDesc &ed = p.descs[ek];
  const SideDesc &sd = ed.s[side];
  const int nt = blockDim.x;
  const int lt = threadIdx.x;
  const int lane = lt & 31;
  const int wl = lt >> 5;
  const int nwarps = nt >> 5;
  const int N = sd.N, K = sd.K, Mout = sd.Mout;
  const int Dnew = ed.Dnew;
  const int ldm = p.ldm;
  const int ncl = (N + CL - 1) / CL;
  const int c0 = rank * ncl;
  const int c_end = min(N, c0 + ncl);

  // shared memory: wleg | rowout | matR | matS | matC | nred
  double *wleg = (double *)smem_raw;
  long long *rowout = (long long *)(wleg + TNSU_MAX_LEGS * p.dcap);
  S *matR = (S *)(rowout + MMAX);
  S *matS = matR + ldm * ldm;
  S *matC = matS + ldm * ldm;
  double *nred = (double *)(matC + ldm * ldm);

  const double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  // NB: lambda_k of THIS edge was just rewritten by K2; only the other legs' weights are used here
  for (int idx = lt; idx < sd.nlegs * p.dcap; idx += nt) {
    int l = idx / p.dcap, x = idx - l * p.dcap;
    wleg[idx] = (x < sd.dims[l]) ? wpoint[(size_t)sd.leg_edge[l] * p.dcap + x] : 1.0;
  }
  for (int r = lt; r < MMAX; r += nt) {
    int s = r / Dnew, x = r - s * Dnew;
    rowout[r] = (r < Mout) ? (long long)s * sd.out_stride_phys + (long long)x * sd.out_stride[sd.bond_leg] : 0;
  }
  const S *small = p.small + (size_t)item * SM_COUNT * ldm * ldm;
  const S *gR = small + (SM_R0 + side) * ldm * ldm;
  const S *gS = small + (SM_S0 + side) * ldm * ldm;
  for (int idx = lt; idx < Mout * ldm; idx += nt) matR[idx] = gR[idx];
  for (int idx = lt; idx < K * ldm; idx += nt) matS[idx] = gS[idx];

  // my columns of Y^H
  const S *ybase = (const S *)sd.ybase + (size_t)b * sd.point_stride;
  S col[NCOLS][MMAX];
  long long out_off[NCOLS];
  double inv_w[NCOLS];
#pragma unroll
  for (int j = 0; j < NCOLS; ++j) {
    const int c = c0 + lt + j * nt;
#pragma unroll
    for (int k = 0; k < MMAX; ++k) col[j][k] = (c < c_end && k < K) ? ybase[(size_t)k * N + c] : from_real<S>(0.0);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NCOLS; ++j) {
    const ColInfo ci = decode_column(sd, (c0 + lt + j * nt < c_end) ? c0 + lt + j * nt : 0, wleg, p.dcap);
    out_off[j] = ci.out_off;
    inv_w[j] = ci.inv_w;
  }
  // C = r~ S  (Mout x K); S = Y1 T^H is lower triangular
  for (int idx = lt; idx < Mout * K; idx += nt) {
    const int i = idx / K, kp = idx - i * K;
    S a = from_real<S>(0.0);
    for (int c = kp; c < K; ++c) fma_acc(a, matR[i * ldm + c], matS[c * ldm + kp]);
    matC[i * ldm + kp] = a;
  }
  __syncthreads();

  // P'[:,c] = [c<K] r~[:,c] - C Yh[:,c]; un-absorb; store in place; accumulate the norm
  S *tbase = (S *)sd.base + (size_t)b * sd.point_stride;
  double n2 = 0.0;
  for (int i = 0; i < Mout; ++i) {
    S o[NCOLS];
#pragma unroll
    for (int j = 0; j < NCOLS; ++j) {
      const int c = c0 + lt + j * nt;
      o[j] = (c < K) ? matR[i * ldm + c] : from_real<S>(0.0);
    }
#pragma unroll
    for (int k = 0; k < MMAX; ++k) {
      if (k < K) {
        const S cik = matC[i * ldm + k];
#pragma unroll
        for (int j = 0; j < NCOLS; ++j) fms_acc(o[j], cik, col[j][k]);
      }
    }
    const long long ro = rowout[i];
#pragma unroll
    for (int j = 0; j < NCOLS; ++j) {
      const int c = c0 + lt + j * nt;
      if (c < c_end) {
        const S v = o[j] * inv_w[j];
        n2 += abs2(v);
        tbase[out_off[j] + ro] = v;
      }
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, m);
  if (lane == 0) nred[wl] = n2;
  __syncthreads();
  if (lt == 0) {
    double tot = 0.0;
    for (int w = 0; w < nwarps; ++w) tot += nred[w];
    const size_t ti = (size_t)b * p.n_tensors + sd.tensor;
    if (CL == 1) {
      p.tscale[ti] = 1.0 / sqrt(tot);
    } else {
      // column slices: the last slice to arrive turns the accumulated norm into the lazy scale
      atomicAdd(&p.tnorm2[ti], tot);
      __threadfence();
      if (atomicAdd(&p.tticket[ti], 1u) == (unsigned)(CL - 1)) {
        __threadfence();
        const double all = atomicAdd(&p.tnorm2[ti], 0.0);
        p.tscale[ti] = 1.0 / sqrt(all);
        p.tnorm2[ti] = 0.0;
        p.tticket[ti] = 0u;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers (instantiated in edge_inst.cu)
// ---------------------------------------------------------------------------------------------
template <int MM, int NC>
cudaError_t launch_lq_rolled(const EdgeLaunch<double> &p, int n_items, cudaStream_t st) {
  if (p.lq_stream_grid > 0) {
    auto kern = lq_rolled_kernel<MM, NC, true, false>;
    if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
    const int grid = (2 * n_items < p.lq_stream_grid) ? 2 * n_items : p.lq_stream_grid;
    kern<<<grid, p.nt, p.smem_lq, st>>>(p);
    return cudaGetLastError();
  }
  if (p.cluster > 1) {
    // one thread-block cluster per (item, side): the CTAs exchange their partial Gram rows through DSMEM
    auto kernc = lq_rolled_kernel<MM, NC, false, true>;
    if (cudaError_t s = tnsu_opt_in_smem((const void *)kernc); s != cudaSuccess) return s;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2u * n_items * p.cluster);
    cfg.blockDim = dim3(p.nt);
    cfg.dynamicSmemBytes = p.smem_lq;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = p.cluster;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernc, p);
  }
  auto kern = lq_rolled_kernel<MM, NC, false, false>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  const int grid = (p.lq_flags != nullptr && 2 * n_items > 1184) ? 1184 : 2 * n_items;  // fallback scan: 8 CTAs per SM
  kern<<<grid, p.nt, p.smem_lq, st>>>(p);
  return cudaGetLastError();
}

template <typename S, int MM, int NC>
cudaError_t launch_lq_variant(const EdgeLaunch<S> &p, int n_items, cudaStream_t st) {
  if constexpr (std::is_same<S, double>::value) {
    if (p.lq_rolled) return launch_lq_rolled<MM, NC>(p, n_items, st);
  }
  if constexpr (!std::is_same<S, double>::value) {
    if (p.cluster > 1) {
      auto kernc = lq_kernel<S, MM, NC, true>;
      if (cudaError_t s = tnsu_opt_in_smem((const void *)kernc); s != cudaSuccess) return s;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2u * n_items * p.cluster);
      cfg.blockDim = dim3(p.nt);
      cfg.dynamicSmemBytes = p.smem_lq;
      cfg.stream = st;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = p.cluster;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      return cudaLaunchKernelEx(&cfg, kernc, p);
    }
  }
  auto kern = lq_kernel<S, MM, NC>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  kern<<<2 * n_items, p.nt, p.smem_lq, st>>>(p);
  return cudaGetLastError();
}

template <typename S, int MM, int NC>
cudaError_t launch_rc_variant(const EdgeLaunch<S> &p, int n_items, cudaStream_t st) {
  auto kern = recompose_kernel<S, MM, NC>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  kern<<<2 * n_items * p.cluster, p.nt, p.smem_rc, st>>>(p);
  return cudaGetLastError();
}

template <typename S, int RPL>
cudaError_t launch_svd_rpl(const EdgeLaunch<S> &p, int n_items, cudaStream_t st) {
  auto kern = svd_kernel<S, RPL>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  kern<<<n_items, SVD_THREADS, p.smem_svd, st>>>(p);
  return cudaGetLastError();
}

// rows of theta per lane -> kernel variant
template <typename S>
cudaError_t launch_svd(const EdgeLaunch<S> &p, int n_items, cudaStream_t st) {
  const int rpl = (p.nr_max + SVD_LPP - 1) / SVD_LPP;
  if (rpl <= 1) return launch_svd_rpl<S, 1>(p, n_items, st);
  if (rpl <= 2) return launch_svd_rpl<S, 2>(p, n_items, st);
  if (rpl <= 3) return launch_svd_rpl<S, 3>(p, n_items, st);
  if (rpl <= 4) return launch_svd_rpl<S, 4>(p, n_items, st);
  if (rpl <= 5) return launch_svd_rpl<S, 5>(p, n_items, st);
  if (rpl <= 6) return launch_svd_rpl<S, 6>(p, n_items, st);
  return launch_svd_rpl<S, 8>(p, n_items, st);
}
