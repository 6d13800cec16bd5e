// lq_chol.cuh — streaming K1 / K3 of the real float64 path: CholeskyQR2 on the FP64 tensor cores.
//
// Replaces, for well-conditioned bond matrices, the RQ reduction scipy.linalg.rq(P) (reference
// src/tnsu/simple_update.py:243-250) and the recompose einsum r~ . Q (:274-275) with its inverse-weight absorption
// (:290-291) and tensor_norm (:294-296).  Any orthonormal row basis Q is admissible for the reduction;
// here P = L Q with L lower triangular from two Cholesky passes:
//
//   pass 1   G1 = P P^T (DMMA Gram, the pair-RDM loop)        G1 = L1 L1^T        W1 = L1^-1
//   pass 2   Q1 = W1 P  (DMMA, P streamed a second time: L2 hits),   G2 = Q1 Q1^T = L2 L2^T      W2 = L2^-1
//            L = L1 L2,   L^-1 = W2 W1
//
// The accumulator fragment of Q1 = W1 P (lane (g,t): rows 8I+g, columns 2t, 2t+1 of an 8-column tile) is, register
// for register, the A AND the B fragment of the Gram update when the k index of the two k-steps is taken as column
// 2t and 2t+1: no shuffle between the two products.  Orthogonality of Q is O(eps) as long as eps*cond(P)^2 << 1; the
// kernel flags an (item, side) whose Cholesky breaks down or whose diag(L1) spans more than 1e5 (cond(P) >~ 1e5), and
// the Householder kernels (lq_rolled_kernel / recompose_dmma_kernel) then process exactly the flagged ones.
//
// K3 needs no reflectors on this route:  P' = r~ Q = (r~ L^-1) P, and the environment weights that were absorbed into
// P are divided out again right away, so  T'[:, c] = tscale * (r~ L^-1) T[:, c]:  one (d D' x d D_k) x (d D_k x N)
// product straight from the OLD tensor into the NEW one, in place column tile by column tile.  The Y^H round trip
// through HBM (half of the edge update's traffic) disappears.
#pragma once
#include "edge_update.cuh"

#define LQC_THREADS 64   // K1c: two warps per (item, side) — the Cholesky phases keep one of them idle, not three
#define RCC_THREADS 32   // K3c: one warp per (item, side) too (4 warps: 4.98 ms, 2 warps: 4.48 ms, 1 warp: 4.40 ms per sweep)
#define LQC_CHUNK 512
#define LQC_WL 32

struct __align__(16) LqcCol {
  double ws;  // product of the environment weights times the lazy tensor scale
  int off;    // byte offset of (s = 0, x_k = 0) of this column
  int pad;
};
struct __align__(16) LqcLeg {
  int dim;
  unsigned magic;
  int stride;
  int pad;
};

DEV void lqc_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

static inline int lqc_smem_bytes(int MB) {
  const int MP = 8 * MB;
  return LQC_CHUNK * 16 + TNSU_MAX_LEGS * 16 + TNSU_MAX_LEGS * LQC_WL * 8 + 5 * MP * (MP + 1) * 8 + 64;
}

// ---- small dense helpers, one warp, lane = row (or column) index, everything in registers -----------------------
// Cholesky of the symmetric MP x MP matrix whose row `lane` is g[] (rows/cols >= M must hold the identity): on
// return l[k] = L[lane][k] (zero above the diagonal), invd = 1 / L[lane][lane]; returns false on breakdown or when
// diag(L) spans more than 1e5.
template <int MP>
DEV bool lqc_cholesky(double (&g)[MP], double (&l)[MP], double &invd, int lane, int M) {
  bool ok = true;
  double dmin = 1e300, dmax = 0.0;
  invd = 1.0;
#pragma unroll
  for (int k = 0; k < MP; ++k) {
    const double d = __shfl_sync(0xffffffffu, g[k], k);
    if (k < M) ok = ok && (d > 0.0) && (d < 1e300);
    const double dd = (d > 0.0) ? d : 1.0;
    const double s = rsqrt(dd);
    const double lkk = dd * s;
    if (k < M) {
      dmin = fmin(dmin, lkk);
      dmax = fmax(dmax, lkk);
    }
    const double lk = (lane >= k) ? g[k] * s : 0.0;
    l[k] = lk;
    if (lane == k) invd = s;
#pragma unroll
    for (int j = k + 1; j < MP; ++j) {
      const double lj = __shfl_sync(0xffffffffu, lk, j);
      g[j] = fma(-lk, lj, g[j]);
    }
  }
  return ok && (dmin > 1e-5 * dmax);
}
// x[r] = (L^-1)[r][lane] (column `lane` of the inverse) by forward substitution; l[] / invd as from lqc_cholesky
template <int MP>
DEV void lqc_tri_inverse(const double (&l)[MP], double invd, double (&x)[MP], int lane) {
#pragma unroll
  for (int r = 0; r < MP; ++r) {
    double a0 = (r == lane) ? 1.0 : 0.0, a1 = 0.0;
#pragma unroll
    for (int k = 0; k < r; ++k) {
      const double lrk = __shfl_sync(0xffffffffu, l[k], r);  // L[r][k] lives in lane r
      if (k & 1)
        a1 = fma(-lrk, x[k], a1);
      else
        a0 = fma(-lrk, x[k], a0);
    }
    x[r] = (a0 + a1) * __shfl_sync(0xffffffffu, invd, r);
  }
}


// ---- complex128 versions (lane = complex row, MC = MP / 2 complex rows): the Gram matrix arrives as the real 2M x 2M Gram
// of the row-interleaved (re, im) matrix, G[r][c] = (g[2r][2c] + g[2r+1][2c+1]) + i (g[2r+1][2c] - g[2r][2c+1]) ----
template <int MC>
DEV bool lqc_cholesky_cx(double (&gre)[MC], double (&gim)[MC], double (&lre)[MC], double (&lim)[MC], double &invd, int lane, int M) {
  bool ok = true;
  double dmin = 1e300, dmax = 0.0;
  invd = 1.0;
#pragma unroll
  for (int k = 0; k < MC; ++k) {
    const double d = __shfl_sync(0xffffffffu, gre[k], k);  // the diagonal of a Hermitian matrix is real
    if (k < M) ok = ok && (d > 0.0) && (d < 1e300);
    const double dd = (d > 0.0) ? d : 1.0;
    const double s = rsqrt(dd);
    const double lkk = dd * s;
    if (k < M) {
      dmin = fmin(dmin, lkk);
      dmax = fmax(dmax, lkk);
    }
    const double lr = (lane > k) ? gre[k] * s : (lane == k ? lkk : 0.0);
    const double li = (lane > k) ? gim[k] * s : 0.0;
    lre[k] = lr;
    lim[k] = li;
    if (lane == k) invd = s;
#pragma unroll
    for (int j = k + 1; j < MC; ++j) {
      // G[lane][j] -= L[lane][k] conj(L[j][k])
      const double jr = __shfl_sync(0xffffffffu, lr, j), ji = __shfl_sync(0xffffffffu, li, j);
      gre[j] = fma(-lr, jr, fma(-li, ji, gre[j]));
      gim[j] = fma(-li, jr, fma(lr, ji, gim[j]));
    }
  }
  return ok && (dmin > 1e-5 * dmax);
}
// x[r] = (L^-1)[r][lane] by forward substitution (complex lower triangular L with a real positive diagonal)
template <int MC>
DEV void lqc_tri_inverse_cx(const double (&lre)[MC], const double (&lim)[MC], double invd, double (&xre)[MC], double (&xim)[MC],
                            int lane) {
#pragma unroll
  for (int r = 0; r < MC; ++r) {
    double ar = (r == lane) ? 1.0 : 0.0, ai = 0.0;
#pragma unroll
    for (int k = 0; k < r; ++k) {
      const double kr = __shfl_sync(0xffffffffu, lre[k], r), ki = __shfl_sync(0xffffffffu, lim[k], r);  // L[r][k] lives in lane r
      ar = fma(-kr, xre[k], fma(ki, xim[k], ar));
      ai = fma(-kr, xim[k], fma(-ki, xre[k], ai));
    }
    const double dinv = __shfl_sync(0xffffffffu, invd, r);
    xre[r] = ar * dinv;
    xim[r] = ai * dinv;
  }
}

// ---- the same two steps for the one-warp-per-matrix kernel with the lane-to-lane broadcasts through shared memory instead
// of shuffles: an FP64 shuffle is ~6 SASS instructions, the O(MP^2) of them made these helpers 10 k of the kernel's 15.8 k
// instructions (27 k for complex128) against a 32 KB instruction cache — ncu (profiles/r02) showed 0.7 (complex: 3.2)
// issue slots lost to instruction fetch per slot used.  `scr` = 64 doubles of per-warp scratch (128 for complex). ----
template <int MP>
DEV bool lqc_cholesky_s(double (&g)[MP], double (&l)[MP], double &invd, int lane, int M, double *scr) {
  bool ok = true;
  double dmin = 1e300, dmax = 0.0;
  invd = 1.0;
#pragma unroll
  for (int k = 0; k < MP; ++k) {
    double *col = scr + (k & 1) * 32;
    const double d = __shfl_sync(0xffffffffu, g[k], k);
    if (k < M) ok = ok && (d > 0.0) && (d < 1e300);
    const double dd = (d > 0.0) ? d : 1.0;
    const double s = rsqrt(dd);
    const double lkk = dd * s;
    if (k < M) {
      dmin = fmin(dmin, lkk);
      dmax = fmax(dmax, lkk);
    }
    const double lk = (lane >= k) ? g[k] * s : 0.0;
    l[k] = lk;
    if (lane == k) invd = s;
    col[lane] = lk;
    __syncwarp();
#pragma unroll
    for (int j = k + 1; j < MP; ++j) g[j] = fma(-lk, col[j], g[j]);
  }
  return ok && (dmin > 1e-5 * dmax);
}
// Ls: L row-major with leading dimension LD, dinv[r] = 1 / L[r][r], both in shared memory and complete
template <int MP>
DEV void lqc_tri_inverse_s(const double *Ls, int LD, const double *dinv, double (&x)[MP], int lane) {
#pragma unroll
  for (int r = 0; r < MP; ++r) {
    double a0 = (r == lane) ? 1.0 : 0.0, a1 = 0.0;
    const double *lr = Ls + r * LD;
#pragma unroll
    for (int k = 0; k < r; ++k) {
      if (k & 1)
        a1 = fma(-lr[k], x[k], a1);
      else
        a0 = fma(-lr[k], x[k], a0);
    }
    x[r] = (a0 + a1) * dinv[r];
  }
}
template <int MC>
DEV bool lqc_cholesky_cx_s(double (&gre)[MC], double (&gim)[MC], double (&lre)[MC], double (&lim)[MC], double &invd, int lane, int M,
                           double *scr) {
  bool ok = true;
  double dmin = 1e300, dmax = 0.0;
  invd = 1.0;
#pragma unroll
  for (int k = 0; k < MC; ++k) {
    double *col = scr + (k & 1) * 64;
    const double d = __shfl_sync(0xffffffffu, gre[k], k);
    if (k < M) ok = ok && (d > 0.0) && (d < 1e300);
    const double dd = (d > 0.0) ? d : 1.0;
    const double s = rsqrt(dd);
    const double lkk = dd * s;
    if (k < M) {
      dmin = fmin(dmin, lkk);
      dmax = fmax(dmax, lkk);
    }
    const double lr = (lane > k) ? gre[k] * s : (lane == k ? lkk : 0.0);
    const double li = (lane > k) ? gim[k] * s : 0.0;
    lre[k] = lr;
    lim[k] = li;
    if (lane == k) invd = s;
    col[2 * lane] = lr;
    col[2 * lane + 1] = li;
    __syncwarp();
#pragma unroll
    for (int j = k + 1; j < MC; ++j) {
      const double jr = col[2 * j], ji = col[2 * j + 1];  // G[lane][j] -= L[lane][k] conj(L[j][k])
      gre[j] = fma(-lr, jr, fma(-li, ji, gre[j]));
      gim[j] = fma(-li, jr, fma(lr, ji, gim[j]));
    }
  }
  return ok && (dmin > 1e-5 * dmax);
}
// Lc: complex L as [MC][MC][2] (re, im) in shared memory, dinv[r] = 1 / L[r][r]
template <int MC>
DEV void lqc_tri_inverse_cx_s(const double *Lc, const double *dinv, double (&xre)[MC], double (&xim)[MC], int lane) {
#pragma unroll
  for (int r = 0; r < MC; ++r) {
    double ar = (r == lane) ? 1.0 : 0.0, ai = 0.0;
    const double *lr = Lc + (size_t)r * MC * 2;
#pragma unroll
    for (int k = 0; k < r; ++k) {
      const double kr = lr[2 * k], ki = lr[2 * k + 1];
      ar = fma(-kr, xre[k], fma(ki, xim[k], ar));
      ai = fma(-kr, xim[k], fma(-ki, xre[k], ai));
    }
    xre[r] = ar * dinv[r];
    xim[r] = ai * dinv[r];
  }
}

// =================================================================================================================
// K1c: one CTA of 4 warps per (item, side)
// =================================================================================================================
template <int MB>
__global__ void __launch_bounds__(LQC_THREADS, MB <= 2 ? 10 : 4) lq_chol_kernel(const EdgeLaunch<double> p, int *flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int MP = 8 * MB, LD = MP + 1, NBLK = MB * (MB + 1) / 2, KS = 2 * MB;
  const int item = blockIdx.x >> 1, side = blockIdx.x & 1;
  const int bi = item / p.n_level_edges;  // position in the launch; the point it stands for comes from the list of active points
  const int ek = p.level_edges[item - bi * p.n_level_edges];
  const int b = p.point_list != nullptr ? p.point_list[bi] : bi;
  if (p.active != nullptr && p.active[b] == 0) return;
  if (p.lqc_force_fallback == 1) {
    if (threadIdx.x == 0) flags[blockIdx.x] = 1;
    return;
  }
  const EdgeDesc &ed = p.descs[ek];
  const SideDesc &sd = ed.s[side];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int M = sd.M, N = sd.N, Dk = ed.Dk, ldm = p.ldm;

  LqcCol *cols = (LqcCol *)smem_raw;                  // [LQC_CHUNK]
  LqcLeg *legs = (LqcLeg *)(cols + LQC_CHUNK);        // environment legs, innermost first
  double *wl = (double *)(legs + TNSU_MAX_LEGS);      // [legs][LQC_WL]
  double *G = wl + TNSU_MAX_LEGS * LQC_WL;            // [MP][LD] Gram matrix of the current pass
  double *L1s = G + MP * LD, *W1s = L1s + MP * LD, *L2s = W1s + MP * LD, *W2s = L2s + MP * LD;
  int *ok_s = (int *)(W2s + MP * LD);

  const double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  const double tsc = p.tscale[(size_t)b * p.n_tensors + sd.tensor];
  for (int idx = tid; idx < TNSU_MAX_LEGS * LQC_WL; idx += LQC_THREADS) {
    const int k = idx / LQC_WL, x = idx - k * LQC_WL;
    int l = sd.nlegs - 1 - k;
    if (l <= sd.bond_leg) --l;
    double w = 1.0;
    if (l >= 0) {
      const int dim = sd.dims[l];
      if (x < dim) w = wpoint[(size_t)sd.leg_edge[l] * p.dcap + x];
      if (x == 0) {
        LqcLeg L;
        L.dim = dim;
        L.magic = dim > 1 ? 0xFFFFFFFFu / (unsigned)dim + 1u : 0u;
        L.stride = (int)sd.in_stride[l];
        L.pad = 0;
        legs[k] = L;
      }
    }
    wl[idx] = w;
  }
  const int nenv = sd.nlegs - 1;
  const char *tbase = (const char *)((const double *)sd.base + (size_t)b * sd.point_stride);
  auto row_ptr = [&](int r) {
    const int rr = r < M ? r : 0;
    const int s = rr / Dk, x = rr - s * Dk;
    return tbase + 8 * ((long long)s * sd.in_stride_phys + (long long)x * sd.in_stride[sd.bond_leg]);
  };
  auto build_table = [&](int c0) {
    for (int cc = tid; cc < LQC_CHUNK; cc += LQC_THREADS) {
      LqcCol e;
      e.ws = 0.0;
      e.off = 0;
      e.pad = 0;
      if (c0 + cc < N) {
        unsigned rem = (unsigned)(c0 + cc);
        double w = tsc;
        unsigned off = 0;
        for (int k = 0; k < nenv; ++k) {
          const LqcLeg L = legs[k];
          const unsigned qq = L.dim > 1 ? __umulhi(rem, L.magic) : rem;
          const unsigned x = rem - qq * (unsigned)L.dim;
          rem = qq;
          off += x * (unsigned)L.stride;
          w *= wl[k * LQC_WL + x];
        }
        e.ws = w;
        e.off = (int)(off * 8u);
      }
      cols[cc] = e;
    }
  };
  // sum the warps' partial Gram fragments in warp order into G (lower blocks), then mirror; pads with the identity
  auto reduce_gram = [&](double (&acc)[NBLK][2]) {
    for (int wv = 0; wv < LQC_THREADS / 32; ++wv) {
      __syncthreads();
      if (warp == wv) {
        int k = 0;
#pragma unroll
        for (int I = 0; I < MB; ++I)
#pragma unroll
          for (int J = 0; J <= I; ++J, ++k) {
            double *d0 = G + (8 * I + g) * LD + 8 * J + 2 * t;
            if (wv == 0) {
              d0[0] = acc[k][0];
              d0[1] = acc[k][1];
            } else {
              d0[0] += acc[k][0];
              d0[1] += acc[k][1];
            }
          }
      }
    }
    __syncthreads();
    for (int idx = tid; idx < MP * MP; idx += LQC_THREADS) {
      const int r = idx / MP, c = idx - r * MP;
      if ((r >> 3) < (c >> 3)) G[r * LD + c] = G[c * LD + r];
    }
    __syncthreads();
  };

  // ------------------------------------------------------------------------------------------------ pass 1: G1 = P P^T
  const char *rowp1[MB];
#pragma unroll
  for (int I = 0; I < MB; ++I) rowp1[I] = row_ptr(8 * I + g);
  double acc[NBLK][2];
#pragma unroll
  for (int k = 0; k < NBLK; ++k) acc[k][0] = acc[k][1] = 0.0;
  for (int c0 = 0; c0 < N; c0 += LQC_CHUNK) {
    __syncthreads();
    build_table(c0);
    __syncthreads();
    const int nks = (min(LQC_CHUNK, N - c0) + 3) >> 2;
#pragma unroll 4
    for (int ks = warp; ks < nks; ks += LQC_THREADS / 32) {
      const LqcCol e = cols[4 * ks + t];
      double v[MB];
#pragma unroll
      for (int I = 0; I < MB; ++I) v[I] = __ldg((const double *)(rowp1[I] + (unsigned)e.off));
      const double w2 = e.ws * e.ws;
      int k = 0;
#pragma unroll
      for (int I = 0; I < MB; ++I) {
        const double a = v[I] * w2;
#pragma unroll
        for (int J = 0; J <= I; ++J, ++k) lqc_dmma(acc[k][0], acc[k][1], a, v[J]);
      }
    }
  }
  reduce_gram(acc);
  // ------------------------------------------------------------------------------------------------ L1, W1 (warp 0)
  double l1[MP];
  if (warp == 0) {
    double gr[MP], invd, x[MP];
#pragma unroll
    for (int j = 0; j < MP; ++j) gr[j] = (lane < M && j < M) ? G[lane * LD + j] : (lane == j ? 1.0 : 0.0);
    const bool ok = lqc_cholesky<MP>(gr, l1, invd, lane, M);
    lqc_tri_inverse<MP>(l1, invd, x, lane);
    if (lane < MP) {
#pragma unroll
      for (int r = 0; r < MP; ++r) {
        W1s[r * LD + lane] = x[r];       // column `lane` of W1
        L1s[lane * LD + r] = l1[r];      // row `lane` of L1
      }
    }
    if (lane == 0) ok_s[0] = ok ? 1 : 0;
  }
  __syncthreads();
  if (ok_s[0] == 0) {  // ill-conditioned: the Householder kernels take this (item, side)
    if (tid == 0) flags[blockIdx.x] = 1;
    return;
  }
  // ------------------------------------------------------------------------------------------------ pass 2: G2 = (W1 P)(W1 P)^T
  double afr[MB][KS];
#pragma unroll
  for (int I = 0; I < MB; ++I)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int i = 8 * I + g, k = 4 * ks + t;
      afr[I][ks] = (i < M && k < M) ? W1s[i * LD + k] : 0.0;
    }
  const char *rowp2[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) rowp2[ks] = row_ptr(4 * ks + t);
#pragma unroll
  for (int k = 0; k < NBLK; ++k) acc[k][0] = acc[k][1] = 0.0;
  for (int c0 = 0; c0 < N; c0 += LQC_CHUNK) {
    if (N > LQC_CHUNK) {  // the table of a single chunk is still valid
      __syncthreads();
      build_table(c0);
      __syncthreads();
    }
    const int ntile = (min(LQC_CHUNK, N - c0) + 7) >> 3;
#pragma unroll 2
    for (int tl = warp; tl < ntile; tl += LQC_THREADS / 32) {
      const unsigned off = (unsigned)cols[8 * tl + g].off;
      double bv[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) bv[ks] = __ldg((const double *)(rowp2[ks] + off));
      const double ws0 = cols[8 * tl + 2 * t].ws, ws1 = cols[8 * tl + 2 * t + 1].ws;
      double q0[MB], q1[MB];
#pragma unroll
      for (int I = 0; I < MB; ++I) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) lqc_dmma(d0, d1, afr[I][ks], bv[ks]);
        q0[I] = d0 * ws0;
        q1[I] = d1 * ws1;
      }
      int k = 0;
#pragma unroll
      for (int I = 0; I < MB; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J, ++k) {
          lqc_dmma(acc[k][0], acc[k][1], q0[I], q0[J]);
          lqc_dmma(acc[k][0], acc[k][1], q1[I], q1[J]);
        }
    }
  }
  reduce_gram(acc);
  // ------------------------------------------------------------------------------------------------ L2, W2, L = L1 L2, L^-1 = W2 W1
  if (warp == 0) {
    double gr[MP], l2[MP], invd, x[MP];
#pragma unroll
    for (int j = 0; j < MP; ++j) gr[j] = (lane < M && j < M) ? G[lane * LD + j] : (lane == j ? 1.0 : 0.0);
    const bool ok = lqc_cholesky<MP>(gr, l2, invd, lane, M);
    lqc_tri_inverse<MP>(l2, invd, x, lane);
    if (lane < MP) {
#pragma unroll
      for (int r = 0; r < MP; ++r) {
        W2s[r * LD + lane] = x[r];
        L2s[lane * LD + r] = l2[r];
      }
    }
    if (lane == 0) ok_s[0] = ok ? 1 : 0;
  }
  __syncthreads();
  // L = L1 L2 (what K2 reads) and L^-1 = W2 W1 (what the streaming K3 reads), all threads, strict upper triangle zeroed
  {
    double *small = p.small + (size_t)item * SM_COUNT * ldm * ldm;
    double *gL = small + (SM_L0 + side) * ldm * ldm;
    double *gW = small + (SM_S0 + side) * ldm * ldm;
    for (int idx = tid; idx < M * M; idx += LQC_THREADS) {
      const int r = idx / M, j = idx - r * M;
      double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
      int k = j;
      for (; k + 1 <= r; k += 2) {
        a0 = fma(L1s[r * LD + k], L2s[k * LD + j], a0);
        a1 = fma(L1s[r * LD + k + 1], L2s[(k + 1) * LD + j], a1);
        c0 = fma(W2s[r * LD + k], W1s[k * LD + j], c0);
        c1 = fma(W2s[r * LD + k + 1], W1s[(k + 1) * LD + j], c1);
      }
      if (k <= r) {
        a0 = fma(L1s[r * LD + k], L2s[k * LD + j], a0);
        c0 = fma(W2s[r * LD + k], W1s[k * LD + j], c0);
      }
      gL[r * ldm + j] = a0 + a1;
      gW[r * ldm + j] = c0 + c1;
    }
    if (tid == 0) flags[blockIdx.x] = ok_s[0] ? 0 : 1;
  }
}

// =================================================================================================================
// K1c, one WARP per (item, side): no block barrier, no idle warp during the Cholesky chains — the four warps of a CTA
// are four independent factorisations whose serial phases hide behind each other's DMMAs
// =================================================================================================================
#define LQW_WARPS 4
#ifndef LQW_MB1_BLOCKS
#define LQW_MB1_BLOCKS 8
#endif
#define LQW_CHUNK 128
#define LQW_WL 16

// complex128 keeps L1 as a packed complex matrix (MP^2 / 2 doubles) instead of an MP x (MP + 1) real one, and its packed
// L2 / W2 take the place of the Gram matrix once that sits in registers
static inline int lqw_smem_bytes_per_warp(int MB, bool cx = false) {
  const int MP = 8 * MB;
  const int mats = cx ? 2 * MP * (MP + 1) + MP * MP / 2 : 4 * MP * (MP + 1);
  return LQW_CHUNK * 16 + TNSU_MAX_LEGS * 16 + TNSU_MAX_LEGS * LQW_WL * 8 + mats * 8 + 160 * 8 + 16;
}
// Warps (= independent factorisations) per CTA.  The wide variants (more than 16 real rows: complex128 at d D > 8, real
// at d D > 16) need 25-38 KB of shared memory and > 200 registers per warp: one-warp CTAs let the SM take as many as fit
// (8 complex / 5 real instead of the 4 of one four-warp CTA).
template <int MB, bool CX>
struct LqwCfg {
  static constexpr int warps = (MB > 2) ? 1 : LQW_WARPS;
  static constexpr int min_blocks = (MB > 2) ? (CX ? 8 : 5) : ((MB == 1 && !CX) ? LQW_MB1_BLOCKS : 4);
};

// CX = true: complex128.  The bond matrix is handled as the REAL matrix of its row-interleaved parts (row 2r = Re P[r],
// row 2r + 1 = Im P[r]: the (re, im) pairs of the tensor storage, 8 bytes apart): its 2M x 2M real Gram holds the four
// real products of the complex Gram, and Q1 = W1 P is the real product with the embedding [[Wr, -Wi], [Wi, Wr]] of W1 —
// exactly the 4 real multiply-adds per complex one, no redundancy, on the same DMMA loops.  Only the two M x M Cholesky /
// inverse steps and the final L = L1 L2, L^-1 = W2 W1 are done in complex arithmetic.  MB counts 8-row blocks of the real
// rows (2M <= 8 MB).
template <int MB, bool CX>
__global__ void __launch_bounds__(32 * LqwCfg<MB, CX>::warps, LqwCfg<MB, CX>::min_blocks)
    lq_chol_warp_kernel(const EdgeLaunch<typename std::conditional<CX, cplx, double>::type> p, int *flags, int smem_per_warp) {
  typedef typename std::conditional<CX, cplx, double>::type S;
  constexpr int ESZ = CX ? 16 : 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int MP = 8 * MB, LD = MP + 1, NBLK = MB * (MB + 1) / 2, KS = 2 * MB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = blockIdx.x * LqwCfg<MB, CX>::warps + warp;  // work item = (item, side)
  const int n_work = 2 * p.batch * p.n_level_edges;
  if (w >= n_work) return;
  const int item = w >> 1, side = w & 1;
  const int bi = item / p.n_level_edges;  // position in the launch; the point it stands for comes from the list of active points
  const int ek = p.level_edges[item - bi * p.n_level_edges];
  const int b = p.point_list != nullptr ? p.point_list[bi] : bi;
  if (p.active != nullptr && p.active[b] == 0) return;
  if (p.lqc_force_fallback == 1) {
    if (lane == 0) flags[w] = 1;
    return;
  }
  const EdgeDesc &ed = p.descs[ek];
  const SideDesc &sd = ed.s[side];
  const int g = lane >> 2, t = lane & 3;
  const int M = sd.M, N = sd.N, Dk = ed.Dk, ldm = p.ldm;
  const int MR = CX ? 2 * M : M;  // real rows

  unsigned char *mine = smem_raw + (size_t)warp * smem_per_warp;
  LqcCol *cols = (LqcCol *)mine;                      // [LQW_CHUNK]
  LqcLeg *legs = (LqcLeg *)(cols + LQW_CHUNK);
  double *wl = (double *)(legs + TNSU_MAX_LEGS);      // [legs][LQW_WL]
  double *G = wl + TNSU_MAX_LEGS * LQW_WL;            // Gram matrix of the current pass, later L2
  double *L1s = G + MP * LD, *W1s = L1s + (CX ? MP * MP / 2 : MP * LD), *W2s = CX ? G : W1s + MP * LD;
  double *scr = W1s + (CX ? 1 : 2) * MP * LD;   // [128] column broadcast of the Cholesky steps
  double *dinv = scr + 128;      // [32] reciprocal diagonal of the current Cholesky factor

  const double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  const double tsc = p.tscale[(size_t)b * p.n_tensors + sd.tensor];
  for (int idx = lane; idx < TNSU_MAX_LEGS * LQW_WL; idx += 32) {
    const int k = idx / LQW_WL, x = idx - k * LQW_WL;
    int l = sd.nlegs - 1 - k;
    if (l <= sd.bond_leg) --l;
    double wv = 1.0;
    if (l >= 0) {
      const int dim = sd.dims[l];
      if (x < dim) wv = wpoint[(size_t)sd.leg_edge[l] * p.dcap + x];
      if (x == 0) {
        LqcLeg L;
        L.dim = dim;
        L.magic = dim > 1 ? 0xFFFFFFFFu / (unsigned)dim + 1u : 0u;
        L.stride = (int)sd.in_stride[l];
        L.pad = 0;
        legs[k] = L;
      }
    }
    wl[idx] = wv;
  }
  __syncwarp();
  const int nenv = sd.nlegs - 1;
  const char *tbase = (const char *)sd.base + (size_t)b * sd.point_stride * ESZ;
  auto row_ptr = [&](int r) {
    const int rq = r < MR ? r : 0;
    const int rr = CX ? (rq >> 1) : rq, part = CX ? (rq & 1) : 0;
    const int s = rr / Dk, x = rr - s * Dk;
    return tbase + ESZ * ((long long)s * sd.in_stride_phys + (long long)x * sd.in_stride[sd.bond_leg]) + 8 * part;
  };
  auto build_table = [&](int c0) {
#pragma unroll
    for (int it = 0; it < LQW_CHUNK / 32; ++it) {
      const int cc = lane + 32 * it;
      LqcCol e;
      e.ws = 0.0;
      e.off = 0;
      e.pad = 0;
      if (c0 + cc < N) {
        unsigned rem = (unsigned)(c0 + cc);
        double wv = tsc;
        unsigned off = 0;
        for (int k = 0; k < nenv; ++k) {
          const LqcLeg L = legs[k];
          const unsigned qq = L.dim > 1 ? __umulhi(rem, L.magic) : rem;
          const unsigned x = rem - qq * (unsigned)L.dim;
          rem = qq;
          off += x * (unsigned)L.stride;
          wv *= wl[k * LQW_WL + x];
        }
        e.ws = wv;
        e.off = (int)(off * (unsigned)ESZ);
      }
      cols[cc] = e;
    }
  };
  // fragments -> G (lower blocks), mirror, all by this warp
  auto store_gram = [&](double (&acc)[NBLK][2]) {
    int k = 0;
#pragma unroll
    for (int I = 0; I < MB; ++I)
#pragma unroll
      for (int J = 0; J <= I; ++J, ++k) {
        double *d0 = G + (8 * I + g) * LD + 8 * J + 2 * t;
        d0[0] = acc[k][0];
        d0[1] = acc[k][1];
      }
    __syncwarp();
    if (MB > 1) {
      for (int idx = lane; idx < MP * MP; idx += 32) {
        const int r = idx / MP, c = idx - r * MP;
        if ((r >> 3) < (c >> 3)) G[r * LD + c] = G[c * LD + r];
      }
      __syncwarp();
    }
  };

  // ------------------------------------------------------------------------------------------------ pass 1
  const char *rowp1[MB];
#pragma unroll
  for (int I = 0; I < MB; ++I) rowp1[I] = row_ptr(8 * I + g);
  double acc[NBLK][2];
#pragma unroll
  for (int k = 0; k < NBLK; ++k) acc[k][0] = acc[k][1] = 0.0;
  for (int c0 = 0; c0 < N; c0 += LQW_CHUNK) {
    __syncwarp();
    build_table(c0);
    __syncwarp();
    // software pipeline: batches of PB k-steps, the loads of batch n+1 are in flight while the DMMAs of batch n issue
    // (the table is zero-padded to the end of the chunk, so a partial last batch just adds zeros)
    constexpr int PB = 8;
    const int nks = (min(LQW_CHUNK, N - c0) + 3) >> 2;
    const int nb = (nks + PB - 1) / PB;
    double v[PB][MB], w2[PB];
#pragma unroll
    for (int u = 0; u < PB; ++u) {
      const LqcCol e = cols[4 * u + t];
      w2[u] = e.ws * e.ws;
#pragma unroll
      for (int I = 0; I < MB; ++I) v[u][I] = __ldg((const double *)(rowp1[I] + (unsigned)e.off));
    }
#pragma unroll 1
    for (int bt = 0; bt < nb; ++bt) {
      double vn[PB][MB], wn[PB];
      const int nxt = (bt + 1 < nb) ? bt + 1 : bt;  // the last batch reloads itself (L1 hits)
#pragma unroll
      for (int u = 0; u < PB; ++u) {
        const LqcCol e = cols[4 * (nxt * PB + u) + t];
        wn[u] = e.ws * e.ws;
#pragma unroll
        for (int I = 0; I < MB; ++I) vn[u][I] = __ldg((const double *)(rowp1[I] + (unsigned)e.off));
      }
#pragma unroll
      for (int u = 0; u < PB; ++u) {
        int k = 0;
#pragma unroll
        for (int I = 0; I < MB; ++I) {
          const double a = v[u][I] * w2[u];
#pragma unroll
          for (int J = 0; J <= I; ++J, ++k) lqc_dmma(acc[k][0], acc[k][1], a, v[u][J]);
        }
      }
#pragma unroll
      for (int u = 0; u < PB; ++u) {
        w2[u] = wn[u];
#pragma unroll
        for (int I = 0; I < MB; ++I) v[u][I] = vn[u][I];
      }
    }
  }
  store_gram(acc);
  // ------------------------------------------------------------------------------------------------ L1, W1
  bool ok;
  constexpr int MC = MP / 2;
  double *L1c = L1s;  // CX: complex L1 as [MC][MC][2] (re, im); W1s keeps the real embedding of W1 for the DMMA fragments
  if constexpr (!CX) {
    double gr[MP], l1[MP], invd, x[MP];
#pragma unroll
    for (int j = 0; j < MP; ++j) gr[j] = (lane < M && j < M) ? G[lane * LD + j] : (lane == j ? 1.0 : 0.0);
    ok = lqc_cholesky_s<MP>(gr, l1, invd, lane, M, scr);
    if (lane < MP) {
#pragma unroll
      for (int r = 0; r < MP; ++r) L1s[lane * LD + r] = l1[r];
      dinv[lane] = invd;
    }
    __syncwarp();
    lqc_tri_inverse_s<MP>(L1s, LD, dinv, x, lane);
    if (lane < MP) {
#pragma unroll
      for (int r = 0; r < MP; ++r) W1s[r * LD + lane] = x[r];
    }
  } else {
    double gre[MC], gim[MC], lre[MC], lim[MC], invd, xre[MC], xim[MC];
    const int r2 = 2 * (lane < MC ? lane : 0);
#pragma unroll
    for (int j = 0; j < MC; ++j) {
      const bool in = lane < M && j < M;
      gre[j] = in ? G[r2 * LD + 2 * j] + G[(r2 + 1) * LD + 2 * j + 1] : (lane == j ? 1.0 : 0.0);
      gim[j] = in ? G[(r2 + 1) * LD + 2 * j] - G[r2 * LD + 2 * j + 1] : 0.0;
    }
    ok = lqc_cholesky_cx_s<MC>(gre, gim, lre, lim, invd, lane, M, scr);
    if (lane < MC) {
#pragma unroll
      for (int r = 0; r < MC; ++r) {
        L1c[(lane * MC + r) * 2] = lre[r];  // row `lane` of L1
        L1c[(lane * MC + r) * 2 + 1] = lim[r];
      }
      dinv[lane] = invd;
    }
    __syncwarp();
    lqc_tri_inverse_cx_s<MC>(L1c, dinv, xre, xim, lane);
    if (lane < MC) {
#pragma unroll
      for (int r = 0; r < MC; ++r) {
        // column `lane` of W1, embedded: rows 2r, 2r+1 x columns 2 lane, 2 lane + 1
        W1s[(2 * r) * LD + 2 * lane] = xre[r];
        W1s[(2 * r) * LD + 2 * lane + 1] = -xim[r];
        W1s[(2 * r + 1) * LD + 2 * lane] = xim[r];
        W1s[(2 * r + 1) * LD + 2 * lane + 1] = xre[r];
      }
    }
  }
  __syncwarp();
  if (!ok) {
    if (lane == 0) flags[w] = 1;
    return;
  }
  // ------------------------------------------------------------------------------------------------ pass 2
  double afr[MB][KS];
#pragma unroll
  for (int I = 0; I < MB; ++I)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int i = 8 * I + g, k = 4 * ks + t;
      afr[I][ks] = (i < MR && k < MR) ? W1s[i * LD + k] : 0.0;
    }
  const char *rowp2[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) rowp2[ks] = row_ptr(4 * ks + t);
#pragma unroll
  for (int k = 0; k < NBLK; ++k) acc[k][0] = acc[k][1] = 0.0;
  for (int c0 = 0; c0 < N; c0 += LQW_CHUNK) {
    if (N > LQW_CHUNK) {
      __syncwarp();
      build_table(c0);
      __syncwarp();
    }
    const int ntile = (min(LQW_CHUNK, N - c0) + 7) >> 3;
    // software pipeline over the tiles: the next tile's B fragments are in flight during this tile's DMMAs
    double bv[KS];
    {
      const unsigned off0 = (unsigned)cols[g].off;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) bv[ks] = __ldg((const double *)(rowp2[ks] + off0));
    }
#pragma unroll 1
    for (int tl = 0; tl < ntile; ++tl) {
      double bn[KS];
      {
        const unsigned offn = (unsigned)cols[8 * (tl + 1 < ntile ? tl + 1 : tl) + g].off;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) bn[ks] = __ldg((const double *)(rowp2[ks] + offn));
      }
      const double ws0 = cols[8 * tl + 2 * t].ws, ws1 = cols[8 * tl + 2 * t + 1].ws;
      double q0[MB], q1[MB];
#pragma unroll
      for (int I = 0; I < MB; ++I) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) lqc_dmma(d0, d1, afr[I][ks], bv[ks]);
        q0[I] = d0 * ws0;
        q1[I] = d1 * ws1;
      }
      int k = 0;
#pragma unroll
      for (int I = 0; I < MB; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J, ++k) {
          lqc_dmma(acc[k][0], acc[k][1], q0[I], q0[J]);
          lqc_dmma(acc[k][0], acc[k][1], q1[I], q1[J]);
        }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) bv[ks] = bn[ks];
    }
  }
  store_gram(acc);
  // ------------------------------------------------------------------------------------------------ L2, W2, products
  if constexpr (!CX) {
    {
      double gr[MP], l2[MP], invd, x[MP];
#pragma unroll
      for (int j = 0; j < MP; ++j) gr[j] = (lane < M && j < M) ? G[lane * LD + j] : (lane == j ? 1.0 : 0.0);
      ok = lqc_cholesky_s<MP>(gr, l2, invd, lane, M, scr);
      __syncwarp();
      if (lane < MP) {
#pragma unroll
        for (int r = 0; r < MP; ++r) G[lane * LD + r] = l2[r];  // L2 takes the Gram matrix's place
        dinv[lane] = invd;
      }
      __syncwarp();
      lqc_tri_inverse_s<MP>(G, LD, dinv, x, lane);
      if (lane < MP) {
#pragma unroll
        for (int r = 0; r < MP; ++r) W2s[r * LD + lane] = x[r];
      }
    }
    __syncwarp();
    const double *L2s = G;
    double *small = (double *)p.small + (size_t)item * SM_COUNT * ldm * ldm;
    double *gL = small + (SM_L0 + side) * ldm * ldm;
    double *gW = small + (SM_S0 + side) * ldm * ldm;
    for (int idx = lane; idx < M * M; idx += 32) {
      const int r = idx / M, j = idx - r * M;
      double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
      int k = j;
      for (; k + 1 <= r; k += 2) {
        a0 = fma(L1s[r * LD + k], L2s[k * LD + j], a0);
        a1 = fma(L1s[r * LD + k + 1], L2s[(k + 1) * LD + j], a1);
        c0 = fma(W2s[r * LD + k], W1s[k * LD + j], c0);
        c1 = fma(W2s[r * LD + k + 1], W1s[(k + 1) * LD + j], c1);
      }
      if (k <= r) {
        a0 = fma(L1s[r * LD + k], L2s[k * LD + j], a0);
        c0 = fma(W2s[r * LD + k], W1s[k * LD + j], c0);
      }
      gL[r * ldm + j] = a0 + a1;
      gW[r * ldm + j] = c0 + c1;
    }
    if (lane == 0) flags[w] = ok ? 0 : 1;
  } else {
    double *L2c = W2s;                    // [MC][MC][2]
    double *W2c = W2s + MC * MC * 2;      // [MC][MC][2]  (MP * LD >= 4 MC^2)
    {
      double gre[MC], gim[MC], lre[MC], lim[MC], invd, xre[MC], xim[MC];
      const int r2 = 2 * (lane < MC ? lane : 0);
#pragma unroll
      for (int j = 0; j < MC; ++j) {
        const bool in = lane < M && j < M;
        gre[j] = in ? G[r2 * LD + 2 * j] + G[(r2 + 1) * LD + 2 * j + 1] : (lane == j ? 1.0 : 0.0);
        gim[j] = in ? G[(r2 + 1) * LD + 2 * j] - G[r2 * LD + 2 * j + 1] : 0.0;
      }
      ok = lqc_cholesky_cx_s<MC>(gre, gim, lre, lim, invd, lane, M, scr);
      __syncwarp();
      if (lane < MC) {
#pragma unroll
        for (int r = 0; r < MC; ++r) {
          L2c[(lane * MC + r) * 2] = lre[r];
          L2c[(lane * MC + r) * 2 + 1] = lim[r];
        }
        dinv[lane] = invd;
      }
      __syncwarp();
      lqc_tri_inverse_cx_s<MC>(L2c, dinv, xre, xim, lane);
      if (lane < MC) {
#pragma unroll
        for (int r = 0; r < MC; ++r) {
          W2c[(r * MC + lane) * 2] = xre[r];  // column `lane` of W2
          W2c[(r * MC + lane) * 2 + 1] = xim[r];
        }
      }
    }
    __syncwarp();
    // L = L1 L2 and L^-1 = W2 W1 in complex arithmetic; W1 is read back from its embedding (re at [2r][2c], im at [2r+1][2c])
    cplx *small = (cplx *)p.small + (size_t)item * SM_COUNT * ldm * ldm;
    cplx *gL = small + (SM_L0 + side) * ldm * ldm;
    cplx *gW = small + (SM_S0 + side) * ldm * ldm;
    for (int idx = lane; idx < M * M; idx += 32) {
      const int r = idx / M, j = idx - r * M;
      double ar = 0.0, ai = 0.0, cr = 0.0, ci = 0.0;
      for (int k = j; k <= r; ++k) {
        const double l1r = L1c[(r * MC + k) * 2], l1i = L1c[(r * MC + k) * 2 + 1];
        const double l2r = L2c[(k * MC + j) * 2], l2i = L2c[(k * MC + j) * 2 + 1];
        ar = fma(l1r, l2r, fma(-l1i, l2i, ar));
        ai = fma(l1r, l2i, fma(l1i, l2r, ai));
        const double w2r = W2c[(r * MC + k) * 2], w2i = W2c[(r * MC + k) * 2 + 1];
        const double w1r = W1s[(2 * k) * LD + 2 * j], w1i = W1s[(2 * k + 1) * LD + 2 * j];
        cr = fma(w2r, w1r, fma(-w2i, w1i, cr));
        ci = fma(w2r, w1i, fma(w2i, w1r, ci));
      }
      gL[r * ldm + j] = mkc(ar, ai);
      gW[r * ldm + j] = mkc(cr, ci);
    }
    if (lane == 0) flags[w] = ok ? 0 : 1;
  }
}

// =================================================================================================================
// K3c: T'[:, c] = tscale (r~ L^-1) T[:, c], in place, one warp (= one CTA) per (item, side)
// =================================================================================================================
// CX = true: complex128 through the same real-embedding as lq_chol_warp_kernel: A = the embedding of r~ L^-1 (2 Mout x 2M),
// B = the row-interleaved (re, im) tile of the old tensor.
template <int MB, bool CX>
__global__ void __launch_bounds__(RCC_THREADS, (MB <= 2 ? 4 : 2) * (128 / RCC_THREADS))
    recompose_chol_kernel(const EdgeLaunch<typename std::conditional<CX, cplx, double>::type> p, const int *flags) {
  typedef typename std::conditional<CX, cplx, double>::type S;
  constexpr int ESZ = CX ? 16 : 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int MP = 8 * MB, KS = 2 * MB;
  const int item = blockIdx.x >> 1, side = blockIdx.x & 1;
  const int bi = item / p.n_level_edges;  // position in the launch; the point it stands for comes from the list of active points
  const int ek = p.level_edges[item - bi * p.n_level_edges];
  const int b = p.point_list != nullptr ? p.point_list[bi] : bi;
  if (p.active != nullptr && p.active[b] == 0) return;
  if (flags[blockIdx.x] != 0) return;  // the Householder recompose owns this one
  const EdgeDesc &ed = p.descs[ek];
  const SideDesc &sd = ed.s[side];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int M = sd.M, N = sd.N, Mout = sd.Mout, Dk = ed.Dk, Dnew = ed.Dnew, ldm = p.ldm;
  const int MR = CX ? 2 * M : M, MoutR = CX ? 2 * Mout : Mout;  // real rows

  int *coff = (int *)smem_raw;                             // [LQC_CHUNK] byte offset of each column (same before / after)
  LqcLeg *legs = (LqcLeg *)(coff + LQC_CHUNK);
  S *matR = (S *)(legs + TNSU_MAX_LEGS);                   // r~ (Mout x M)
  S *matW = matR + ldm * ldm;                              // L^-1 (M x M)
  S *matC = matW + ldm * ldm;                              // r~ L^-1
  double *nred = (double *)(matC + ldm * ldm);

  if (tid < TNSU_MAX_LEGS) {
    int l = sd.nlegs - 1 - tid;
    if (l <= sd.bond_leg) --l;
    if (l >= 0) {
      LqcLeg L;
      L.dim = sd.dims[l];
      L.magic = L.dim > 1 ? 0xFFFFFFFFu / (unsigned)L.dim + 1u : 0u;
      L.stride = (int)sd.in_stride[l];
      L.pad = 0;
      legs[tid] = L;
    }
  }
  const S *small = p.small + (size_t)item * SM_COUNT * ldm * ldm;
  const S *gR = small + (SM_R0 + side) * ldm * ldm;
  const S *gW = small + (SM_S0 + side) * ldm * ldm;
  for (int idx = tid; idx < Mout * ldm; idx += RCC_THREADS) matR[idx] = gR[idx];
  for (int idx = tid; idx < M * ldm; idx += RCC_THREADS) matW[idx] = gW[idx];
  const double tsc = p.tscale[(size_t)b * p.n_tensors + sd.tensor];
  __syncthreads();
  for (int idx = tid; idx < Mout * M; idx += RCC_THREADS) {
    const int i = idx / M, j = idx - i * M;
    S a0 = from_real<S>(0.0), a1 = from_real<S>(0.0);
    int k = j;
    for (; k + 1 < M; k += 2) {
      fma_acc(a0, matR[i * ldm + k], matW[k * ldm + j]);
      fma_acc(a1, matR[i * ldm + k + 1], matW[(k + 1) * ldm + j]);
    }
    if (k < M) fma_acc(a0, matR[i * ldm + k], matW[k * ldm + j]);
    matC[i * ldm + j] = (a0 + a1) * tsc;
  }
  __syncthreads();
  double afr[MB][KS];
#pragma unroll
  for (int I = 0; I < MB; ++I)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int i = 8 * I + g, k = 4 * ks + t;
      double v = 0.0;
      if (i < MoutR && k < MR) {
        if constexpr (CX) {
          const cplx cv = ((const cplx *)matC)[(i >> 1) * ldm + (k >> 1)];
          v = ((i & 1) == (k & 1)) ? cv.x : ((i & 1) ? cv.y : -cv.y);  // embedding [[cr, -ci], [ci, cr]]
        } else {
          v = ((const double *)matC)[i * ldm + k];
        }
      }
      afr[I][ks] = v;
    }
  char *tbase = (char *)sd.base + (size_t)b * sd.point_stride * ESZ;
  const char *rowin[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int rq = 4 * ks + t < MR ? 4 * ks + t : 0;
    const int r = CX ? (rq >> 1) : rq, part = CX ? (rq & 1) : 0;
    const int s = r / Dk, x = r - s * Dk;
    rowin[ks] = tbase + ESZ * ((long long)s * sd.in_stride_phys + (long long)x * sd.in_stride[sd.bond_leg]) + 8 * part;
  }
  long long rowout[MB];
#pragma unroll
  for (int I = 0; I < MB; ++I) {
    const int rq = 8 * I + g < MoutR ? 8 * I + g : 0;
    const int r = CX ? (rq >> 1) : rq, part = CX ? (rq & 1) : 0;
    const int s = r / Dnew, x = r - s * Dnew;
    rowout[I] = ESZ * ((long long)s * sd.out_stride_phys + (long long)x * sd.out_stride[sd.bond_leg]) + 8 * part;
  }
  const int nenv = sd.nlegs - 1;
  // 32-byte accesses are possible when the innermost tensor axis is an environment leg whose dimension is a multiple of 4
  // (four consecutive columns are then contiguous) and every other offset keeps the 32-byte alignment.  (The analogous
  // row-vector fetch for an innermost bond leg was measured: no gain, 128 registers.)
  bool vec_cols = false;
  {
    const int lin = sd.nlegs - 1;
    if (!CX && sd.bond_leg != lin && lin >= 0 && sd.dims[lin] % 4 == 0 && sd.in_stride[lin] == 1 && sd.in_stride[sd.bond_leg] % 4 == 0 &&
        sd.in_stride_phys % 4 == 0 && sd.point_stride % 4 == 0 && (((size_t)sd.base & 31) == 0))
      vec_cols = true;
  }
  double n2 = 0.0;
  for (int c0 = 0; c0 < N; c0 += LQC_CHUNK) {
    __syncthreads();
    for (int cc = tid; cc < LQC_CHUNK; cc += RCC_THREADS) {
      unsigned off = 0;
      if (c0 + cc < N) {
        unsigned rem = (unsigned)(c0 + cc);
        for (int k = 0; k < nenv; ++k) {
          const LqcLeg L = legs[k];
          const unsigned qq = L.dim > 1 ? __umulhi(rem, L.magic) : rem;
          off += (rem - qq * (unsigned)L.dim) * (unsigned)L.stride;
          rem = qq;
        }
      }
      coff[cc] = (int)(off * (unsigned)ESZ);
    }
    __syncthreads();
    const int ncols = min(LQC_CHUNK, N - c0);
    const int ntile = (ncols + 7) >> 3;
    int tile0 = 0;
    if (vec_cols) {
      // 32 columns per warp trip with 32-byte accesses: lane (g, t) loads the columns 4g .. 4g+3 of its row with one
      // LDG.256 per k-step; its j-th value is the B fragment of "tile j" = the columns {4 g' + j}.  The accumulators of the
      // four tiles then hold, for this lane, the columns 8t + j and 8t + 4 + j: two contiguous groups of four -> one
      // STG.256 each.  4x fewer load and store instructions than the tile-by-tile loop below.
      const int ngrp = ncols >> 5;
      for (int grp = warp; grp < ngrp; grp += RCC_THREADS / 32) {
        const int cb = 32 * grp;
        const unsigned off_in = (unsigned)coff[cb + 4 * g];
        double bv[KS][4];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                       : "=d"(bv[ks][0]), "=d"(bv[ks][1]), "=d"(bv[ks][2]), "=d"(bv[ks][3])
                       : "l"(rowin[ks] + off_in)
                       : "memory");
        const unsigned oa = (unsigned)coff[cb + 8 * t], ob = (unsigned)coff[cb + 8 * t + 4];
#pragma unroll
        for (int I = 0; I < MB; ++I) {
          double da[4], db[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            da[j] = db[j] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) lqc_dmma(da[j], db[j], afr[I][ks], bv[ks][j]);
          }
          if (8 * I + g < MoutR) {
#pragma unroll
            for (int j = 0; j < 4; ++j) n2 = fma(da[j], da[j], fma(db[j], db[j], n2));
            asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(tbase + rowout[I] + oa), "d"(da[0]), "d"(da[1]),
                         "d"(da[2]), "d"(da[3])
                         : "memory");
            asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(tbase + rowout[I] + ob), "d"(db[0]), "d"(db[1]),
                         "d"(db[2]), "d"(db[3])
                         : "memory");
          }
        }
      }
      tile0 = 4 * ngrp;  // the remaining (< 32) columns of the chunk go tile by tile
    }
#pragma unroll 2
    for (int tl = tile0 + warp; tl < ntile; tl += RCC_THREADS / 32) {
      // the whole (M x 8) input tile is in registers before the first element of the output tile is stored
      const unsigned off = (unsigned)coff[min(8 * tl + g, ncols - 1)];
      double bv[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) bv[ks] = *(const volatile double *)(rowin[ks] + off);
      const int ca = 8 * tl + 2 * t;
      const unsigned o0 = (unsigned)coff[min(ca, ncols - 1)], o1 = (unsigned)coff[min(ca + 1, ncols - 1)];
      double d0[MB], d1[MB];
#pragma unroll
      for (int I = 0; I < MB; ++I) {
        d0[I] = d1[I] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) lqc_dmma(d0[I], d1[I], afr[I][ks], bv[ks]);
      }
      __syncwarp();  // every lane of the warp has its inputs: the stores below overwrite the same tile
#pragma unroll
      for (int I = 0; I < MB; ++I) {
        if (8 * I + g < MoutR) {
          if (ca < ncols) {
            n2 = fma(d0[I], d0[I], n2);
            *(double *)(tbase + rowout[I] + o0) = d0[I];
          }
          if (ca + 1 < ncols) {
            n2 = fma(d1[I], d1[I], n2);
            *(double *)(tbase + rowout[I] + o1) = d1[I];
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, m);
  if (lane == 0) nred[warp] = n2;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < RCC_THREADS / 32; ++w) tot += nred[w];
    p.tscale[(size_t)b * p.n_tensors + sd.tensor] = 1.0 / sqrt(tot);
  }
}

static inline int rcc_smem_bytes(int ldm, int esz = 8) { return LQC_CHUNK * 4 + TNSU_MAX_LEGS * 16 + 3 * ldm * ldm * esz + 64; }

template <int MB>
cudaError_t launch_lq_chol_v(const EdgeLaunch<double> &p, int n_items, int *flags, cudaStream_t st) {
  auto kern = lq_chol_kernel<MB>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  if (p.lqc_force_fallback != 2) {  // default: one warp per (item, side); 2 selects the two-warp-CTA kernel (A/B comparisons)
    auto kw = lq_chol_warp_kernel<MB, false>;
    if (cudaError_t s = tnsu_opt_in_smem((const void *)kw); s != cudaSuccess) return s;
    const int per_warp = lqw_smem_bytes_per_warp(MB);
    constexpr int NW = LqwCfg<MB, false>::warps;
    kw<<<(2 * n_items + NW - 1) / NW, 32 * NW, per_warp * NW, st>>>(p, flags, per_warp);
    return cudaGetLastError();
  }
  kern<<<2 * n_items, LQC_THREADS, lqc_smem_bytes(MB), st>>>(p, flags);
  return cudaGetLastError();
}
template <int MB>
cudaError_t launch_rc_chol_v(const EdgeLaunch<double> &p, int n_items, const int *flags, cudaStream_t st) {
  auto kern = recompose_chol_kernel<MB, false>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  kern<<<2 * n_items, RCC_THREADS, rcc_smem_bytes(p.ldm), st>>>(p, flags);
  return cudaGetLastError();
}
// complex128: m_max complex rows -> 2 m_max real rows
template <int MB>
cudaError_t launch_lq_chol_cx_v(const EdgeLaunch<cplx> &p, int n_items, int *flags, cudaStream_t st) {
  auto kw = lq_chol_warp_kernel<MB, true>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kw); s != cudaSuccess) return s;
  const int per_warp = lqw_smem_bytes_per_warp(MB, true);
  constexpr int NW = LqwCfg<MB, true>::warps;
  kw<<<(2 * n_items + NW - 1) / NW, 32 * NW, per_warp * NW, st>>>(p, flags, per_warp);
  return cudaGetLastError();
}
template <int MB>
cudaError_t launch_rc_chol_cx_v(const EdgeLaunch<cplx> &p, int n_items, const int *flags, cudaStream_t st) {
  auto kern = recompose_chol_kernel<MB, true>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  kern<<<2 * n_items, RCC_THREADS, rcc_smem_bytes(p.ldm, 16), st>>>(p, flags);
  return cudaGetLastError();
}
inline cudaError_t launch_lq_chol_cx(const EdgeLaunch<cplx> &p, int n_items, int m_max, int *flags, cudaStream_t st) {
  const int mb = (2 * m_max + 7) / 8;
  if (mb <= 1) return launch_lq_chol_cx_v<1>(p, n_items, flags, st);
  if (mb == 2) return launch_lq_chol_cx_v<2>(p, n_items, flags, st);
  if (mb == 3) return launch_lq_chol_cx_v<3>(p, n_items, flags, st);
  return launch_lq_chol_cx_v<4>(p, n_items, flags, st);
}
inline cudaError_t launch_rc_chol_cx(const EdgeLaunch<cplx> &p, int n_items, int m_max, const int *flags, cudaStream_t st) {
  const int mb = (2 * m_max + 7) / 8;
  if (mb <= 1) return launch_rc_chol_cx_v<1>(p, n_items, flags, st);
  if (mb == 2) return launch_rc_chol_cx_v<2>(p, n_items, flags, st);
  if (mb == 3) return launch_rc_chol_cx_v<3>(p, n_items, flags, st);
  return launch_rc_chol_cx_v<4>(p, n_items, flags, st);
}
inline cudaError_t launch_lq_chol(const EdgeLaunch<double> &p, int n_items, int m_max, int *flags, cudaStream_t st) {
  const int mb = (m_max + 7) / 8;
  if (mb == 1) return launch_lq_chol_v<1>(p, n_items, flags, st);
  if (mb == 2) return launch_lq_chol_v<2>(p, n_items, flags, st);
  if (mb == 3) return launch_lq_chol_v<3>(p, n_items, flags, st);
  return launch_lq_chol_v<4>(p, n_items, flags, st);
}
inline cudaError_t launch_rc_chol(const EdgeLaunch<double> &p, int n_items, int m_max, const int *flags, cudaStream_t st) {
  const int mb = (m_max + 7) / 8;
  if (mb == 1) return launch_rc_chol_v<1>(p, n_items, flags, st);
  if (mb == 2) return launch_rc_chol_v<2>(p, n_items, flags, st);
  if (mb == 3) return launch_rc_chol_v<3>(p, n_items, flags, st);
  return launch_rc_chol_v<4>(p, n_items, flags, st);
}
