// edge_update.cuh — the fused per-edge imaginary-time-evolution step of Simple Update.
//
// Replaces one iteration of the `for ek in range(m)` loop of SimpleUpdate.simple_update
// (reference src/tnsu/simple_update.py:219-296): absorb weights (tensor_network.py:242-259), permute +
// flatten (:340-369), RQ (:243-250), gate + theta (:456-485), truncated SVD (:393-409), recompose
// (:268-283), un-absorb (tensor_network.py:261-278), normalise and store (:294-296).
//
// One work item = (point b, edge e).  It runs either in ONE CTA whose two halves ("sides") own the two
// site tensors, or in a 2-CTA thread-block cluster (one site tensor per SM, small factors exchanged
// through distributed shared memory).  Within a side, thread t owns NCOLS whole COLUMNS of the absorbed
// bond matrix P (M = d*D_k rows, N = prod(other bond dims) columns) IN REGISTERS for the entire update:
//
//   1. load:   col[r] = T[s, x_k, others] * prod_{m != k} lambda_m[x_m] * tscale      (r = s*D_k + x_k)
//   2. LQ:     K = min(M,N) Householder steps on the ROWS of P.  Step k needs one block reduction
//              g_i = sum_{c>k} P[i,c] conj(P[k,c]) over all rows i: rows i>k give the reflector dot
//              products, rows i<k give u_i^H u_k for the compact-WY factor.  Reflector k stays in row k
//              of the register-resident columns: nothing of size M x N ever touches shared memory.
//   3. theta:  L_i, L_j (M x K, shared memory) -> theta = L_i diag(lambda_k) L_j G; one-sided Jacobi SVD.
//   4. WY:     P' = r~ Q = [r~ 0] - (r~ Y1 T^H) Y^H   — one small GEMM per column, no reductions.
//   5. store:  T' = P' * prod lambda_m^-1 written in place (all reads of T happened in step 1);
//              ||T'||^2 is block-reduced and kept as a lazy scale (tscale) that the next load applies.
//
// Gauge: any orthonormal row basis Q is admissible (SURVEY.md §8a "Gauge facts"); weights, RDMs and
// energies are gauge invariant and are what the parity tests compare.
#pragma once
#include <cooperative_groups.h>

#include <type_traits>

#include "common.cuh"

namespace cg = cooperative_groups;

// compile-time loop: f(std::integral_constant<int, I>) for I in [0, N) — guarantees constant indices into
// register-resident arrays where `#pragma unroll` on a large body is only a hint
template <int I, int N, typename F>
DEV void static_for(F &&f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

struct EdgeSmem {  // byte offsets into dynamic shared memory
  int side_bytes;  // size of one side block
  int o_wleg, o_rowin, o_rowout, o_red, o_colk, o_mat_a, o_mat_b, o_mat_c, o_mat_d, o_tau, o_nred;
  int o_shared;    // start of the shared block (after the side block(s))
  int o_svd_a, o_svd_v, o_sig, o_order, o_lam_old, o_lam_new, o_flag, o_gate, o_peer_l, o_peer_rt;
  int total_bytes;
};

template <typename S>
struct EdgeLaunch {
  const EdgeDesc *descs;     // indexed by edge id
  const int *level_edges;    // edge ids handled by this launch
  int n_level_edges;
  int batch, m_edges, dcap, d4;
  double *weights;           // [batch][m][dcap]
  double *tscale;            // [batch][n_tensors] lazy 1/||T||
  int n_tensors;
  const S *gates;            // [slot][batch][m][d4]
  long long gate_slot_stride;
  const int *point_slot;     // per-point slot (run loop) or nullptr
  int fixed_slot;
  const int *active;         // per-point flag or nullptr
  double *edge_err;          // [batch][m]
  int nt_side;               // threads per side (multiple of 32)
  int ldm;                   // leading dimension of the small matrices
  int lda, ldv;              // leading dimensions of the SVD work arrays (column-major)
  EdgeSmem sm;
  // debug dump (tnsu_debug_edge)
  S *dbg_li, *dbg_lj;
  double *dbg_sigma;
  int *dbg_info;
  int dbg_point;
};

// ---------------------------------------------------------------------------------------------
// warp reduce-scatter of NV per-lane values: afterwards lane l holds the warp total of value (l % NV)
// for NV <= 32 (all lanes with the same l % NV hold the same total).
// ---------------------------------------------------------------------------------------------
template <typename S, int NV>
DEV S warp_reduce_scatter(S (&v)[NV], int lane) {
  static_assert(NV == 1 || NV == 2 || NV == 4 || NV == 8 || NV == 16 || NV == 32, "NV must be a power of two <= 32");
  // Stage with xor-mask `m` halves the live values: a lane keeps the half selected by bit m of its id.
  // Order of bits: the lowest log2(NV) bits of the lane id select the value index.
#pragma unroll
  for (int half = NV / 2; half >= 1; half /= 2) {
    // live values are v[0 .. 2*half); lanes with bit `half` set keep the upper half
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      S keep = upper ? v[j + half] : v[j];
      S send = upper ? v[j] : v[j + half];
      v[j] = keep + shfl_xor(send, half);
    }
  }
  // v[0] now holds the partial total of value (lane % NV) over the lanes that share lane % NV' bits
  // consumed so far; finish over the remaining lane bits.
#pragma unroll
  for (int m = NV; m < 32; m *= 2) v[0] = v[0] + shfl_xor(v[0], m);
  return v[0];
}

// ---------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi SVD in shared memory.  A is nr x nc column-major (nr >= nc), V nc x nc.
// On exit A V_in = A_out with mutually orthogonal columns; V holds the accumulated rotations.
// All threads of the CTA call this (uses __syncthreads).  8 lanes cooperate on one column pair.
// ---------------------------------------------------------------------------------------------
DEV double rot_sign(double x) { return x < 0.0 ? -1.0 : 1.0; }

template <typename S>
__device__ int jacobi_svd(S *A, int lda, int nr, int nc, S *V, int ldv, volatile int *flag, int tid, int nthr) {
  for (int idx = tid; idx < nc * nc; idx += nthr) {
    int col = idx / nc, row = idx - col * nc;
    V[col * ldv + row] = from_real<S>(row == col ? 1.0 : 0.0);
  }
  const int lane8 = tid & 7;
  const int group = tid >> 3;
  const int ngroups = nthr >> 3;
  const unsigned gmask = 0xFFu << (8 * ((tid & 31) >> 3));
  const int np = (nc + 1) & ~1;  // even number of players; index >= nc is a bye
  const int npairs = np / 2;
  const int nrounds = np - 1;
  const double tol2 = 2.25e-30;  // (1.5e-15)^2
  int sweeps = 0;
  __syncthreads();
  for (; sweeps < 40; ++sweeps) {
    if (tid == 0) *flag = 0;
    __syncthreads();
    for (int r = 0; r < nrounds; ++r) {
      for (int pp = group; pp < npairs; pp += ngroups) {
        int p, q;
        if (pp == 0) {
          p = np - 1;
          q = r;
        } else {
          p = (r + pp) % (np - 1);
          q = (r - pp + (np - 1)) % (np - 1);
        }
        if (p > q) {
          int t = p;
          p = q;
          q = t;
        }
        if (q < nc) {
          S *ap = A + (size_t)p * lda;
          S *aq = A + (size_t)q * lda;
          double al = 0.0, be = 0.0;
          S ga = from_real<S>(0.0);
          for (int i = lane8; i < nr; i += 8) {
            S x = ap[i], y = aq[i];
            al += abs2(x);
            be += abs2(y);
            fma_acc(ga, cj(x), y);
          }
#pragma unroll
          for (int m = 4; m >= 1; m >>= 1) {
            al += __shfl_xor_sync(gmask, al, m);
            be += __shfl_xor_sync(gmask, be, m);
            ga = ga + shfl_xor(ga, m, gmask);
          }
          const double g2 = abs2(ga);
          if (g2 > tol2 * al * be && g2 > 0.0) {
            const double absg = sqrt(g2);
            const S ph = ga * (1.0 / absg);  // e^{i phi}
            const double delta = be - al;
            const double inv_r = rsqrt(delta * delta + 4.0 * g2);
            const double c2 = 0.5 + 0.5 * fabs(delta) * inv_r;
            const double inv_c = rsqrt(c2);
            const double c = c2 * inv_c;
            const double s = rot_sign(delta) * absg * inv_r * inv_c;
            const S sp = cj(ph) * s;  // s e^{-i phi}
            const S sq = ph * s;      // s e^{+i phi}
            for (int i = lane8; i < nr; i += 8) {
              S x = ap[i], y = aq[i];
              ap[i] = x * c - sp * y;
              aq[i] = sq * x + y * c;
            }
            S *vp = V + (size_t)p * ldv;
            S *vq = V + (size_t)q * ldv;
            for (int i = lane8; i < nc; i += 8) {
              S x = vp[i], y = vq[i];
              vp[i] = x * c - sp * y;
              vq[i] = sq * x + y * c;
            }
            if (lane8 == 0) *flag = 1;
          }
        }
      }
      __syncthreads();
    }
    if (*flag == 0) {
      ++sweeps;
      break;
    }
    __syncthreads();  // everyone has read the flag before the next sweep resets it
  }
  return sweeps;
}

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
template <typename S, int MMAX, int NCOLS>
constexpr int edge_max_threads() {
  return (MMAX * NCOLS * (int)sizeof(S) / 8 > 48) ? 256 : 512;
}

template <typename S, int MMAX, int NCOLS, bool CLUSTER>
__global__ void __launch_bounds__(edge_max_threads<S, MMAX, NCOLS>(), 1) edge_update_kernel(const EdgeLaunch<S> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];

  const int work = CLUSTER ? (blockIdx.x >> 1) : blockIdx.x;
  const int b = work / p.n_level_edges;
  const int ek = p.level_edges[work - b * p.n_level_edges];
  if (p.active != nullptr && p.active[b] == 0) return;  // uniform over the CTA (and the cluster)

  const EdgeDesc &ed = p.descs[ek];
  const int nt = p.nt_side;
  const int side = CLUSTER ? (int)(blockIdx.x & 1) : (int)(threadIdx.x / nt);
  const int lt = CLUSTER ? (int)threadIdx.x : (int)(threadIdx.x - side * nt);
  const int lane = lt & 31;
  const int wl = lt >> 5;
  const int nwarps = nt >> 5;
  const SideDesc &sd = ed.s[side];
  const int M = sd.M, N = sd.N, K = sd.K, Mout = sd.Mout;
  const int d = ed.d, Dk = ed.Dk, Dnew = ed.Dnew;
  const int ldm = p.ldm;
  constexpr int CH = MMAX < 8 ? MMAX : 8;

  // ---- shared memory views ----
  unsigned char *sblk = smem_raw + (CLUSTER ? 0 : side * p.sm.side_bytes);
  double *wleg = (double *)(sblk + p.sm.o_wleg);          // [nlegs][dcap]
  long long *rowin = (long long *)(sblk + p.sm.o_rowin);  // [MMAX]
  long long *rowout = (long long *)(sblk + p.sm.o_rowout);
  S *red = (S *)(sblk + p.sm.o_red);                      // [2][TNSU_MAX_WARPS][MMAX]
  S *colk = (S *)(sblk + p.sm.o_colk);                    // [2][MMAX]
  S *matL = (S *)(sblk + p.sm.o_mat_a);                   // L (M x K), later C (Mout x K)
  S *matZ = (S *)(sblk + p.sm.o_mat_b);                   // Z (K x K), later Sm = Y1 T^H (K x K)
  S *matT = (S *)(sblk + p.sm.o_mat_c);                   // T (K x K), later r~ (Mout x K)
  S *matY = (S *)(sblk + p.sm.o_mat_d);                   // first K columns of Y^H (K x K)
  double *tau_s = (double *)(sblk + p.sm.o_tau);          // [MMAX]
  double *nred = (double *)(sblk + p.sm.o_nred);          // [TNSU_MAX_WARPS]
  unsigned char *shb = smem_raw + p.sm.o_shared;
  S *svdA = (S *)(shb + p.sm.o_svd_a);
  S *svdV = (S *)(shb + p.sm.o_svd_v);
  double *sig = (double *)(shb + p.sm.o_sig);
  int *order = (int *)(shb + p.sm.o_order);
  double *lam_old = (double *)(shb + p.sm.o_lam_old);
  double *lam_new = (double *)(shb + p.sm.o_lam_new);
  int *flag = (int *)(shb + p.sm.o_flag);
  S *gate_s = (S *)(shb + p.sm.o_gate);
  S *peerL = (S *)(shb + p.sm.o_peer_l);    // cluster: rank 0's copy of L_j
  S *peerRt = (S *)(shb + p.sm.o_peer_rt);  // cluster: r~_j staged in rank 0 for rank 1

  auto side_sync = [&]() {
    if (CLUSTER)
      __syncthreads();
    else
      bar_sync(1 + side, nt);
  };

  double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  long long clk[8];
  clk[0] = clock64();

  // ---- 0. weights of every leg, row offsets, old lambda_k ----
  for (int idx = lt; idx < sd.nlegs * p.dcap; idx += nt) {
    int l = idx / p.dcap, x = idx - l * p.dcap;
    wleg[idx] = (x < sd.dims[l]) ? wpoint[(size_t)sd.leg_edge[l] * p.dcap + x] : 1.0;
  }
  for (int r = lt; r < MMAX; r += nt) {
    int s = r / Dk, x = r - s * Dk;
    rowin[r] = (r < M) ? (long long)s * sd.in_stride_phys + (long long)x * sd.in_stride[sd.bond_leg] : 0;
  }
  for (int idx = lt; idx < ldm * ldm; idx += nt) matL[idx] = from_real<S>(0.0);
  if (side == 0 || CLUSTER)
    for (int x = lt; x < Dk; x += nt) lam_old[x] = wpoint[(size_t)ek * p.dcap + x];
  side_sync();

  // ---- 1. load my columns, absorbing the environment weights ----
  S *tbase = (S *)sd.base + (size_t)b * sd.point_stride;
  const double tsc = p.tscale[(size_t)b * p.n_tensors + sd.tensor];
  S col[NCOLS][MMAX];
  long long out_off[NCOLS];
  double inv_w[NCOLS];
#pragma unroll
  for (int j = 0; j < NCOLS; ++j) {
    const int c = lt + j * nt;
    long long ioff = 0, ooff = 0;
    double w = 1.0, iw = 1.0;
    int rem = c;
    for (int l = sd.nlegs - 1; l >= 0; --l) {
      if (l == sd.bond_leg) continue;
      const int dim = sd.dims[l];
      const int x = rem % dim;
      rem /= dim;
      ioff += (long long)x * sd.in_stride[l];
      ooff += (long long)x * sd.out_stride[l];
      const double wl_x = wleg[l * p.dcap + x];
      w *= wl_x;
      iw *= 1.0 / wl_x;
    }
    out_off[j] = ooff;
    inv_w[j] = iw;
    const bool valid = c < N;
    const double wc = w * tsc;
#pragma unroll
    for (int r = 0; r < MMAX; ++r) {
      S v = from_real<S>(0.0);
      if (valid && r < M) v = tbase[ioff + rowin[r]] * wc;
      col[j][r] = v;
    }
  }

  clk[1] = clock64();
  // ---- 2. Householder LQ on the rows of P; reflectors stay in col[][k] ----
  static_for<0, MMAX>([&](auto kk_c) {
    constexpr int kk = decltype(kk_c)::value;
    if (kk < K) {
      constexpr int buf = kk & 1;
      // partial Gram row g_i = sum_{c>kk, my columns} P[i,c] conj(P[kk,c]), rows in chunks of CH to bound
      // live registers; warp reduce-scatter leaves row (cb + lane) in lane < CH
#pragma unroll
      for (int j = 0; j < NCOLS; ++j) {
        if (lt + j * nt == kk) {
#pragma unroll
          for (int i = 0; i < MMAX; ++i) colk[buf * MMAX + i] = col[j][i];
        }
      }
#pragma unroll
      for (int cb = 0; cb < MMAX; cb += CH) {
        if (cb < M) {
          S acc[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) acc[i] = from_real<S>(0.0);
#pragma unroll
          for (int j = 0; j < NCOLS; ++j) {
            if (lt + j * nt > kk) {
              const S pk = cj(col[j][kk]);
#pragma unroll
              for (int i = 0; i < CH; ++i) fma_acc(acc[i], col[j][cb + i], pk);
            }
          }
          S tot = warp_reduce_scatter<S, CH>(acc, lane);
          if (lane < CH) red[(buf * TNSU_MAX_WARPS + wl) * MMAX + cb + lane] = tot;
        }
      }
      side_sync();

      // every warp finalises redundantly: lane l owns row l
      S g = from_real<S>(0.0);
      S mycol = from_real<S>(0.0);
      if (lane < M) {
        for (int w = 0; w < nwarps; ++w) g = g + red[(buf * TNSU_MAX_WARPS + w) * MMAX + lane];
        mycol = colk[buf * MMAX + lane];
      }
      const S alpha = colk[buf * MMAX + kk];
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
          const int c = lt + j * nt;
          if (c > kk) fms_acc(col[j][i], wi, col[j][kk]);
        }
      }
#pragma unroll
      for (int j = 0; j < NCOLS; ++j) {
        const int c = lt + j * nt;
        if (c == kk)
          col[j][kk] = head;
        else if (c < kk)
          col[j][kk] = from_real<S>(0.0);
      }
    }
  });

  clk[2] = clock64();
  // first K columns of Y^H to shared memory: matY[k][c] = conj(u_k[c])
#pragma unroll
  for (int j = 0; j < NCOLS; ++j) {
    const int c = lt + j * nt;
    if (c < K) {
#pragma unroll
      for (int k = 0; k < MMAX; ++k)
        if (k < K) matY[k * ldm + c] = col[j][k];
    }
  }
  side_sync();

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
  side_sync();
  // Sm[c][k'] = (Y1 T^H)[c][k'] = conj( sum_k Yh[k][c] T[k'][k] ), overwrites Z
  for (int idx = lt; idx < K * K; idx += nt) {
    const int c = idx / K, kp = idx - c * K;
    S a = from_real<S>(0.0);
    for (int k = kp; k <= c; ++k) fma_acc(a, matY[k * ldm + c], matT[kp * ldm + k]);
    matZ[c * ldm + kp] = cj(a);
  }

  clk[3] = clock64();
  // ---- 3. exchange L, theta, SVD ----
  if constexpr (CLUSTER)
    cg::this_cluster().sync();
  else
    __syncthreads();

  const bool svd_owner = CLUSTER ? (side == 0) : true;
  const int ctid = threadIdx.x;
  const int cthreads = blockDim.x;
  const SideDesc &si = ed.s[0];
  const SideDesc &sj = ed.s[1];
  const int Ki = si.K, Kj = sj.K;
  const int ni = ed.ni, nj = ed.nj;
  const bool transposed = ni < nj;  // Jacobi wants rows >= cols
  const int nr = transposed ? nj : ni;
  const int nc = transposed ? ni : nj;
  S *rt_i = nullptr, *rt_j = nullptr;
  const S *Li = nullptr, *Lj = nullptr;
  if (CLUSTER) {
    Li = matL;
    Lj = peerL;
    rt_i = matT;
    rt_j = peerRt;
  } else {
    Li = (S *)(smem_raw + p.sm.o_mat_a);
    Lj = (S *)(smem_raw + p.sm.side_bytes + p.sm.o_mat_a);
    rt_i = (S *)(smem_raw + p.sm.o_mat_c);
    rt_j = (S *)(smem_raw + p.sm.side_bytes + p.sm.o_mat_c);
  }
  int svd_sweeps = 0;
  clk[4] = clk[5] = clk[3];
  if (svd_owner) {
    if constexpr (CLUSTER) {
      const S *remote = cg::this_cluster().map_shared_rank(matL, 1);
      for (int idx = ctid; idx < sj.M * ldm; idx += cthreads) peerL[idx] = remote[idx];
    }
    const int slot = p.point_slot ? p.point_slot[b] : p.fixed_slot;
    const S *gsrc = p.gates + (size_t)slot * p.gate_slot_stride + ((size_t)b * p.m_edges + ek) * p.d4;
    for (int idx = ctid; idx < p.d4; idx += cthreads) gate_s[idx] = gsrc[idx];
    __syncthreads();

    // theta[(qi,i'),(j',qj)] = sum_{i,j,e} L_i[(i,e),qi] lam_e L_j[(j,e),qj] G[i,j,i',j']
    for (int idx = ctid; idx < Ki * Kj; idx += cthreads) {
      const int qi = idx / Kj, qj = idx - qi * Kj;
      S x[TNSU_MAX_SPIN][TNSU_MAX_SPIN];
#pragma unroll
      for (int i = 0; i < TNSU_MAX_SPIN; ++i)
#pragma unroll
        for (int j = 0; j < TNSU_MAX_SPIN; ++j) {
          S a = from_real<S>(0.0);
          if (i < d && j < d)
            for (int e = 0; e < Dk; ++e)
              fma_acc(a, Li[(i * Dk + e) * ldm + qi] * lam_old[e], Lj[(j * Dk + e) * ldm + qj]);
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
              svdA[(size_t)row * p.lda + cl] = cj(a);  // A = theta^H : A[cl][row]
            else
              svdA[(size_t)cl * p.lda + row] = a;
          }
        }
    }
    clk[4] = clock64();
    svd_sweeps = jacobi_svd<S>(svdA, p.lda, nr, nc, svdV, p.ldv, flag, ctid, cthreads);
    clk[5] = clock64();

    // singular values, descending order
    for (int j = ctid; j < nc; j += cthreads) {
      double s2 = 0.0;
      for (int i = 0; i < nr; ++i) s2 += abs2(svdA[(size_t)j * p.lda + i]);
      sig[j] = sqrt(s2);
    }
    __syncthreads();
    for (int j = ctid; j < nc; j += cthreads) {
      const double sj_ = sig[j];
      int pos = 0;
      for (int i = 0; i < nc; ++i) {
        const double si_ = sig[i];
        pos += (si_ > sj_ || (si_ == sj_ && i < j)) ? 1 : 0;
      }
      order[pos] = j;
    }
    __syncthreads();
    if (ctid == 0) {
      double tot = 0.0;
      for (int x = 0; x < Dnew; ++x) tot += sig[order[x]];
      double err2 = 0.0;
      for (int x = 0; x < Dnew; ++x) {
        const double ln = sig[order[x]] / tot;
        lam_new[x] = ln;
        wpoint[(size_t)ek * p.dcap + x] = ln;
        const double df = ln - lam_old[x];
        err2 += df * df;
      }
      p.edge_err[(size_t)b * p.m_edges + ek] = (Dnew == Dk) ? sqrt(err2) : nan("");
    }
    // r~_i[(i',x'),qi] = U[(qi,i'),x'],   r~_j[(j',x'),qj] = V^H[x',(j',qj)]
    for (int idx = ctid; idx < d * Dnew * Ki; idx += cthreads) {
      const int row = idx / Ki, qi = idx - row * Ki;
      const int ip = row / Dnew, xp = row - ip * Dnew;
      const int oc = order[xp];
      const double sv = sig[oc];
      const double isv = sv > 0.0 ? 1.0 / sv : 0.0;
      const int trow = qi * d + ip;
      rt_i[row * ldm + qi] = transposed ? svdV[(size_t)oc * p.ldv + trow] : svdA[(size_t)oc * p.lda + trow] * isv;
    }
    for (int idx = ctid; idx < d * Dnew * Kj; idx += cthreads) {
      const int row = idx / Kj, qj = idx - row * Kj;
      const int jp = row / Dnew, xp = row - jp * Dnew;
      const int oc = order[xp];
      const double sv = sig[oc];
      const double isv = sv > 0.0 ? 1.0 / sv : 0.0;
      const int tcol = jp * Kj + qj;
      rt_j[row * ldm + qj] =
          transposed ? cj(svdA[(size_t)oc * p.lda + tcol] * isv) : cj(svdV[(size_t)oc * p.ldv + tcol]);
    }
    if (p.dbg_info != nullptr && b == p.dbg_point) {
      for (int idx = ctid; idx < si.M * Ki; idx += cthreads) p.dbg_li[idx] = Li[(idx / Ki) * ldm + idx % Ki];
      for (int idx = ctid; idx < sj.M * Kj; idx += cthreads) p.dbg_lj[idx] = Lj[(idx / Kj) * ldm + idx % Kj];
      for (int j = ctid; j < nc; j += cthreads) p.dbg_sigma[j] = sig[order[j]];
      if (ctid == 0) {
        p.dbg_info[0] = si.M;
        p.dbg_info[1] = Ki;
        p.dbg_info[2] = sj.M;
        p.dbg_info[3] = Kj;
        p.dbg_info[4] = ni;
        p.dbg_info[5] = nj;
        p.dbg_info[6] = Dnew;
        p.dbg_info[7] = svd_sweeps;
      }
    }
  }
  if constexpr (CLUSTER) {
    cg::this_cluster().sync();
    if (side == 1) {
      const S *remote = cg::this_cluster().map_shared_rank(peerRt, 0);
      for (int idx = ctid; idx < Mout * ldm; idx += cthreads) matT[idx] = remote[idx];
    }
    cg::this_cluster().sync();  // rank 0's shared memory must outlive rank 1's remote reads
  } else {
    __syncthreads();
  }

  clk[6] = clock64();
  // ---- 4. C = r~ (Y1 T^H)  (Mout x K), overwrites L ----
  for (int idx = lt; idx < Mout * K; idx += nt) {
    const int i = idx / K, kp = idx - i * K;
    S a = from_real<S>(0.0);
    for (int c = 0; c < K; ++c) fma_acc(a, matT[i * ldm + c], matZ[c * ldm + kp]);
    matL[i * ldm + kp] = a;
  }
  for (int r = lt; r < MMAX; r += nt) {
    int s = r / Dnew, x = r - s * Dnew;
    rowout[r] = (r < Mout) ? (long long)s * sd.out_stride_phys + (long long)x * sd.out_stride[sd.bond_leg] : 0;
  }
  side_sync();

  // ---- 5. P'[:,c] = [c<K] r~[:,c] - C Yh[:,c]; un-absorb; store in place; accumulate the norm ----
  double n2 = 0.0;
  for (int i = 0; i < Mout; ++i) {
    S o[NCOLS];
#pragma unroll
    for (int j = 0; j < NCOLS; ++j) {
      const int c = lt + j * nt;
      o[j] = (c < K) ? matT[i * ldm + c] : from_real<S>(0.0);
    }
#pragma unroll
    for (int k = 0; k < MMAX; ++k) {
      if (k < K) {
        const S cik = matL[i * ldm + k];
#pragma unroll
        for (int j = 0; j < NCOLS; ++j) fms_acc(o[j], cik, col[j][k]);
      }
    }
    const long long ro = rowout[i];
#pragma unroll
    for (int j = 0; j < NCOLS; ++j) {
      const int c = lt + j * nt;
      if (c < N) {
        const S v = o[j] * inv_w[j];
        n2 += abs2(v);
        tbase[out_off[j] + ro] = v;
      }
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, m);
  if (lane == 0) nred[wl] = n2;
  side_sync();
  if (lt == 0) {
    double tot = 0.0;
    for (int w = 0; w < nwarps; ++w) tot += nred[w];
    p.tscale[(size_t)b * p.n_tensors + sd.tensor] = 1.0 / sqrt(tot);
  }
  if (p.dbg_info != nullptr && b == p.dbg_point && threadIdx.x == 0 && side == 0) {
    clk[7] = clock64();
    // phase cycles: load | LQ | WY factors | theta | SVD | extract+exchange | recompose+store
    for (int i = 0; i < 7; ++i) p.dbg_info[8 + i] = (int)(clk[i + 1] - clk[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// host launcher of one kernel variant (instantiated in edge_inst.cu)
// ---------------------------------------------------------------------------------------------
#define TNSU_SMEM_LIMIT (227 * 1024)
template <typename S, int MM, int NC, bool CL>
cudaError_t launch_edge_variant(const EdgeLaunch<S> &p, int n_work, int threads, int smem, cudaStream_t st) {
  auto kern = edge_update_kernel<S, MM, NC, CL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t s = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TNSU_SMEM_LIMIT);
    if (s != cudaSuccess) return s;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CL ? 2 * n_work : n_work);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (CL) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, p);
}

