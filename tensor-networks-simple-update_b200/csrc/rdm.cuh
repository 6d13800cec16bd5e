// rdm.cuh — reduced-density-matrix kernels behind energy_per_site and the expectation family.
//
// Replaces SimpleUpdate.tensor_pair_rdm (reference src/tnsu/simple_update.py:525-593, the ncon call at
// :584-589), tensor_pair_expectation (:605-613), tensor_rdm (:498-523) and the per-edge / per-site
// sums of energy_per_site (:615-637), pair_expectation_per_site (:639-649), expectation_per_site
// (:651-661).
//
// The ncon network of tensor_pair_rdm is Gram(T_i) and Gram(T_j) — every leg except the shared edge
// is traced inside its own tensor with its weight absorbed — joined by diag(lambda_k) on the ket and
// on the bra side.  One CTA per (point, edge) streams each site tensor ONCE through a column-chunk
// staging buffer, accumulates the Hermitian M x M Gram matrix (M = d*D_k) with one register
// accumulator per (row,row') pair, and finishes with the tiny d^4 contraction.
#pragma once
#include "common.cuh"

#define RDM_THREADS 256
#define RDM_CHUNK 64  // columns staged per pass

template <typename S>
struct PairRdmLaunch {
  const EdgeDesc *descs;
  const int *edge_list;  // edges handled by this launch
  int first_edge;        // >= 0: the list is first_edge, first_edge + 1, ... (saves a dependent global load per CTA)
  int n_list;
  int batch, m_edges, dcap, d4;
  const int *point_list;         // nullptr, or the points this launch covers (batch = its length)
  const double *weights;
  const S *ops;                  // operator per (point, edge) [batch][m][d4], or one shared op [d4]
  long long op_point_stride;     // 0 when the operator is shared by all points
  long long op_edge_stride;      // 0 when shared by all edges
  cplx *pair_val;                // [batch][m]  sum rho*O per (point, edge); may be null
  cplx *rdm_out;                 // [batch][d4] trace-normalised rho for edge_list[0]; may be null
  const int *iter_state;         // run loop: per-point iteration index (check only when even and > 0) or null
  const int *active;             // per-point flag or null
};

template <typename S>
__global__ void __launch_bounds__(RDM_THREADS) pair_rdm_kernel(const PairRdmLaunch<S> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int bi = blockIdx.x / p.n_list;
  const int ek = p.edge_list[blockIdx.x - bi * p.n_list];
  const int b = p.point_list != nullptr ? p.point_list[bi] : bi;
  if (p.active != nullptr && p.active[b] == 0) return;
  if (p.iter_state != nullptr) {
    const int it = p.iter_state[b];
    if (!(it % 2 == 0 && it > 0)) return;
  }
  const EdgeDesc &ed = p.descs[ek];
  const int tid = threadIdx.x;
  const int d = ed.d, Dk = ed.Dk;
  const int M = d * Dk;  // both sides share d and D_k
  const int ldc = RDM_CHUNK + 1;

  // shared memory: gram[2][M*M] | chunk[M][ldc] | coloff[CHUNK] | colw[CHUNK] | wleg[legs][dcap] | lam | rho | red
  S *gram = (S *)smem_raw;
  S *chunk = gram + 2 * M * M;
  long long *coloff = (long long *)(chunk + M * ldc);
  double *colw = (double *)(coloff + RDM_CHUNK);
  double *wleg = colw + RDM_CHUNK;
  double *lam = wleg + TNSU_MAX_LEGS * p.dcap;
  cplx *rho = (cplx *)(lam + p.dcap);
  cplx *red = rho + p.d4;

  const double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  for (int x = tid; x < Dk; x += RDM_THREADS) lam[x] = wpoint[(size_t)ek * p.dcap + x];

  const int npairs = M * (M + 1) / 2;
  for (int side = 0; side < 2; ++side) {
    const SideDesc &sd = ed.s[side];
    const S *tbase = (const S *)sd.base + (size_t)b * sd.point_stride;
    __syncthreads();
    for (int idx = tid; idx < sd.nlegs * p.dcap; idx += RDM_THREADS) {
      int l = idx / p.dcap, x = idx - l * p.dcap;
      wleg[idx] = (x < sd.dims[l]) ? wpoint[(size_t)sd.leg_edge[l] * p.dcap + x] : 1.0;
    }
    // (row,row') pairs owned by this thread: pair index q -> r >= r'
    int pr[3], pc[3];
    S acc[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      int q = tid + u * RDM_THREADS;
      acc[u] = from_real<S>(0.0);
      pr[u] = -1;
      pc[u] = 0;
      if (q < npairs) {
        int r = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
        while (r * (r + 1) / 2 > q) --r;
        while ((r + 1) * (r + 2) / 2 <= q) ++r;
        pr[u] = r;
        pc[u] = q - r * (r + 1) / 2;
      }
    }
    for (int c0 = 0; c0 < sd.N; c0 += RDM_CHUNK) {
      __syncthreads();
      if (tid < RDM_CHUNK) {
        const int c = c0 + tid;
        long long off = 0;
        double w = 0.0;
        if (c < sd.N) {
          w = 1.0;
          int rem = c;
          for (int l = sd.nlegs - 1; l >= 0; --l) {
            if (l == sd.bond_leg) continue;
            const int dim = sd.dims[l];
            const int x = rem % dim;
            rem /= dim;
            off += (long long)x * sd.in_stride[l];
            w *= wleg[l * p.dcap + x];
          }
        }
        coloff[tid] = off;
        colw[tid] = w;
      }
      __syncthreads();
      for (int idx = tid; idx < M * RDM_CHUNK; idx += RDM_THREADS) {
        const int r = idx / RDM_CHUNK, cc = idx - r * RDM_CHUNK;
        const int s = r / Dk, x = r - s * Dk;
        S v = from_real<S>(0.0);
        if (c0 + cc < sd.N)
          v = tbase[coloff[cc] + (long long)s * sd.in_stride_phys + (long long)x * sd.in_stride[sd.bond_leg]] *
              colw[cc];
        chunk[r * ldc + cc] = v;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        if (pr[u] >= 0) {
          const S *a = chunk + pr[u] * ldc;
          const S *bb = chunk + pc[u] * ldc;
          S t = acc[u];
          for (int cc = 0; cc < RDM_CHUNK; ++cc) fma_acc(t, a[cc], cj(bb[cc]));
          acc[u] = t;
        }
      }
    }
    S *g = gram + side * M * M;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      if (pr[u] >= 0) {
        g[pr[u] * M + pc[u]] = acc[u];
        if (pr[u] != pc[u]) g[pc[u] * M + pr[u]] = cj(acc[u]);
      }
    }
  }
  __syncthreads();

  // rho[i,j,i',j'] = sum_{e,e'} lam_e lam_e' G_i[(i,e),(i',e')] G_j[(j,e),(j',e')]
  const S *gi = gram, *gj = gram + M * M;
  for (int idx = tid; idx < p.d4; idx += RDM_THREADS) {
    int jp = idx % d, ip = (idx / d) % d, j = (idx / (d * d)) % d, i = idx / (d * d * d);
    S a = from_real<S>(0.0);
    for (int e = 0; e < Dk; ++e)
      for (int f = 0; f < Dk; ++f)
        fma_acc(a, gi[(i * Dk + e) * M + ip * Dk + f] * (lam[e] * lam[f]), gj[(j * Dk + e) * M + jp * Dk + f]);
    rho[idx] = to_c128(a);
  }
  __syncthreads();
  if (tid == 0) {
    cplx tr = mkc(0.0, 0.0);
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) tr += rho[((i * d + j) * d + i) * d + j];
    red[0] = tr;
  }
  __syncthreads();
  {
    const cplx tr = red[0];
    const double den = 1.0 / abs2(tr);
    const cplx itr = mkc(tr.x * den, -tr.y * den);
    __syncthreads();
    for (int idx = tid; idx < p.d4; idx += RDM_THREADS) rho[idx] = rho[idx] * itr;
    __syncthreads();
    if (p.rdm_out != nullptr && ek == p.edge_list[0])
      for (int idx = tid; idx < p.d4; idx += RDM_THREADS) p.rdm_out[(size_t)b * p.d4 + idx] = rho[idx];
    if (p.pair_val != nullptr && tid == 0) {
      const S *op = p.ops + (size_t)b * p.op_point_stride + (size_t)ek * p.op_edge_stride;
      cplx v = mkc(0.0, 0.0);
      for (int idx = 0; idx < p.d4; ++idx) v += rho[idx] * to_c128(op[idx]);
      p.pair_val[(size_t)b * p.m_edges + ek] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Real-FP64 fast path: the two Gram matrices on the FP64 tensor cores (DMMA m8n8k4), operands loaded
// straight from HBM/L2 into the mma fragment layout — no staging of the tensor in shared memory.
//
// G[r][r'] = sum_c P[r][c] w_c P[r'][c] is a SYRK with M = d*D_k <= 32 rows and N = prod(other dims)
// columns.  For m8n8k4 the A fragment (row-major 8x4: lane holds A[lane/4][lane%4]) and the B fragment
// (col-major 4x8: lane holds B[lane%4][lane/4]) of a Gram product are the SAME register when the row
// block is the same, so one load per (8-row block, 4 columns) feeds every tile of that block: MB loads
// and MB(MB+1)/2 DMMAs per k-step (the strictly upper blocks follow by symmetry).  One warp per tensor
// (PG_WPS = 1: warp 0 takes tensor i, warp 1 tensor j — measured 16 % faster at D = 8 and 1.85x at D = 4
// than four warps per side, whose partial fragments had to meet in shared memory through four barrier
// rounds; with PG_WPS > 1 the partial sums are still added in a fixed order).  Column offsets and the
// squared environment weights come from a per-chunk table (128 columns) in shared memory.
// CTAs of one point are adjacent in the grid, so a site tensor shared by several edges is read from
// DRAM once and from L2 afterwards.
// ---------------------------------------------------------------------------------------------
#define PG_WPS 1      // warps per side: one warp owns a whole Gram matrix (no partial sums, one barrier per chunk)
#define PG_THREADS (64 * PG_WPS)
#define PG_CHUNK 128  // columns per side and pass
#define PG_UNROLL 4   // k-steps in flight per warp
#define PG_WL 32      // leading dimension of the per-leg weight table (D <= 16 whenever d*D <= 32)

DEV void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 16- / 32-byte read-only global loads (LDG.E.128 / LDG.E.256 on sm_100a); the address must be aligned to the width
template <int W>
DEV void pg_ldvec(const char *ptr, double (&v)[W]) {
  static_assert(W == 2 || W == 4, "vector width");
  if constexpr (W == 2)
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(ptr));
  else
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(ptr));
}

struct __align__(16) PgCol {  // one column of the bond matrix: squared environment weight and element offset
  double w;
  int off;
  int pad;
};
struct __align__(16) PgLeg {  // one environment leg: dimension, 2^32/dim rounded up (exact division of c < 2^20), stride
  int dim;
  unsigned magic;
  int stride;
  int pad;
};

static inline int pair_dmma_smem_bytes(int MB, int d4) {
  const int MP = 8 * MB;
  return 2 * MP * (MP + 1) * 8 + 2 * PG_CHUNK * 16 + 2 * TNSU_MAX_LEGS * 16 + 2 * TNSU_MAX_LEGS * PG_WL * 8 + MP * 8 + d4 * 8 + 64;
}

template <int MB>
__global__ void __launch_bounds__(PG_THREADS, (MB <= 2 ? 4 : 2) * (256 / PG_THREADS)) pair_rdm_dmma_kernel(const PairRdmLaunch<double> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int MP = 8 * MB;
  constexpr int NBLK = MB * (MB + 1) / 2;
  const int bi = blockIdx.x / p.n_list;
  const int ek = p.first_edge >= 0 ? p.first_edge + (int)(blockIdx.x - bi * p.n_list) : p.edge_list[blockIdx.x - bi * p.n_list];
  const int b = p.point_list != nullptr ? p.point_list[bi] : bi;
  const int ek_first = p.first_edge >= 0 ? p.first_edge : p.edge_list[0];
  if (p.active != nullptr && p.active[b] == 0) return;
  if (p.iter_state != nullptr) {
    const int it = p.iter_state[b];
    if (!(it % 2 == 0 && it > 0)) return;
  }
  const EdgeDesc &ed = p.descs[ek];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int side = warp / PG_WPS, q = warp % PG_WPS, g = lane >> 2, t = lane & 3;
  const int d = ed.d, Dk = ed.Dk, M = d * Dk;

  PgCol *cols = (PgCol *)smem_raw;                           // [2][PG_CHUNK]
  PgLeg *legs = (PgLeg *)(cols + 2 * PG_CHUNK);              // [2][legs]: environment legs, innermost first
  double *gram = (double *)(legs + 2 * TNSU_MAX_LEGS);       // [2][MP][MP]
  constexpr int GLD = MP + 1;  // odd leading dimension of the two Gram matrices: the mirror store and the (e, f) reads of the
                               // contraction were 8-way / 2-way bank conflicts with GLD = MP (ncu r02: 34.5 M of 92.8 M wavefronts)
  double *wsq = gram + 2 * MP * GLD;                          // [2][legs][PG_WL] squared weights, same order as `legs`
  double *lamrow = wsq + 2 * TNSU_MAX_LEGS * PG_WL;          // [MP] lambda_k of the bond index of row r
  double *rho = lamrow + MP;                                 // [d4]
  double *red = rho + p.d4;

  const double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  // environment legs of both sides, innermost first (the order in which a column index is decoded)
  for (int idx = tid; idx < 2 * TNSU_MAX_LEGS * PG_WL; idx += PG_THREADS) {
    const int sdx = idx / (TNSU_MAX_LEGS * PG_WL), k = (idx / PG_WL) % TNSU_MAX_LEGS, x = idx % PG_WL;
    const SideDesc &s2 = ed.s[sdx];
    int l = s2.nlegs - 1 - k;          // k-th environment leg from the inside: skip the bond leg
    if (l <= s2.bond_leg) --l;
    double w = 1.0;
    if (l >= 0) {
      const int dim = s2.dims[l];
      if (x < dim) w = wpoint[(size_t)s2.leg_edge[l] * p.dcap + x];
      if (x == 0) {
        PgLeg L;
        L.dim = dim;
        L.magic = dim > 1 ? 0xFFFFFFFFu / (unsigned)dim + 1u : 0u;
        L.stride = (int)s2.in_stride[l];
        L.pad = 0;
        legs[sdx * TNSU_MAX_LEGS + k] = L;
      }
    }
    wsq[idx] = w * w;
  }
  if (tid < MP) lamrow[tid] = tid < M ? wpoint[(size_t)ek * p.dcap + tid % Dk] : 0.0;

  const SideDesc &sd = ed.s[side];
  // pull both site tensors towards L2 while the column table is being built (each is one contiguous block)
  {
    const char *tb = (const char *)((const double *)sd.base + (size_t)b * sd.point_stride);
    const long long bytes = (long long)M * sd.N * 8;
    for (long long o = (long long)(tid % (32 * PG_WPS)) * 128; o < bytes; o += 32 * PG_WPS * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(tb + o));
  }
  // rows past M are clamped to row 0: their Gram entries are computed and never read
  // How the fragments are fetched (warp-uniform, per side).  The LSU wavefront rate, not DRAM, bounded the first
  // version (8 lines touched per 8-byte-per-lane load), so wherever the layout allows it one lane fetches 32 bytes:
  //   mode 1  the innermost tensor axis is an environment leg (dim % 4 == 0): a lane takes FOUR consecutive columns of
  //           its row with one LDG.256 and feeds four k-steps from it.  Within a 16-column group k-step j then holds
  //           the columns 4t + j instead of 4j + t — a permutation of the summation index, the same for A and B.
  //   mode 2  the bond leg is the innermost axis: the rows (s, x_k .. x_k + MB - 1) are contiguous, so fragment row g of
  //           block I is bond-matrix row g * MB + I and ONE 16/32-byte load serves all MB row blocks of a k-step.
  //   mode 0  everything else (odd dimensions, unaligned slices): one 8-byte load per row block and k-step.
  int mode = 0;
  {
    const bool inner_bond = sd.bond_leg == sd.nlegs - 1;
    const int lin = sd.nlegs - 1 - (inner_bond ? 1 : 0);  // innermost environment leg
    const bool aligned = (((size_t)sd.base & 31) == 0) && (sd.point_stride % 4 == 0) && (sd.in_stride_phys % 4 == 0);
    if (!inner_bond) {
      if (aligned && lin >= 0 && sd.dims[lin] % 4 == 0 && sd.in_stride[lin] == 1 && sd.in_stride[sd.bond_leg] % 4 == 0) mode = 1;
    } else if ((MB == 2 || MB == 4) && sd.nlegs >= 2) {
      if (aligned && sd.in_stride[sd.bond_leg] == 1 && Dk % MB == 0 && (Dk % 4 == 0 || MB == 2)) mode = 2;
    }
  }
  const char *rowp[MB];
#pragma unroll
  for (int I = 0; I < MB; ++I) {
    const int rr = (mode == 2) ? g * MB + I : 8 * I + g;
    const int r = rr < M ? rr : 0;
    const int s = r / Dk, x = r - s * Dk;
    rowp[I] = (const char *)((const double *)sd.base + (size_t)b * sd.point_stride + (long long)s * sd.in_stride_phys +
                             (long long)x * sd.in_stride[sd.bond_leg]);
  }
  double acc[NBLK][2];
#pragma unroll
  for (int k = 0; k < NBLK; ++k) acc[k][0] = acc[k][1] = 0.0;

  const int nenv0 = ed.s[0].nlegs - 1, nenv1 = ed.s[1].nlegs - 1;
  const int N0 = ed.s[0].N, N1 = ed.s[1].N;
  const int Nmax = max(N0, N1);
  constexpr int KSTEP_GROUP = PG_WPS * PG_UNROLL;  // k-steps per trip of the warps of a side
  for (int c0 = 0; c0 < Nmax; c0 += PG_CHUNK) {
    __syncthreads();  // legs / wsq ready, previous chunk consumed
    // each thread decodes 2 * PG_CHUNK / PG_THREADS entries, unrolled: the weight products are dependent FP64 chains
#pragma unroll
    for (int it = 0; it < 2 * PG_CHUNK / PG_THREADS; ++it) {
      const int idx = tid + it * PG_THREADS;
      const int sdx = idx / PG_CHUNK, cc = idx - sdx * PG_CHUNK;  // uniform per warp
      const int Ns = sdx ? N1 : N0;
      const int left = Ns - c0;
      const int nks_pad = left > 0 ? (((min(left, PG_CHUNK) + 3) >> 2) + KSTEP_GROUP - 1) / KSTEP_GROUP * KSTEP_GROUP : 0;
      if (cc < 4 * nks_pad) {
        PgCol e;
        e.w = 0.0;
        e.off = 0;
        e.pad = 0;
        if (cc < left) {
          const PgLeg *lg = legs + sdx * TNSU_MAX_LEGS;
          const double *wl = wsq + sdx * TNSU_MAX_LEGS * PG_WL;
          const int nenv = sdx ? nenv1 : nenv0;
          unsigned rem = (unsigned)(c0 + cc);
          double w = 1.0;
          unsigned off = 0;
          for (int k = 0; k < nenv; ++k) {
            const PgLeg L = lg[k];
            const unsigned qq = L.dim > 1 ? __umulhi(rem, L.magic) : rem;
            const unsigned x = rem - qq * (unsigned)L.dim;
            rem = qq;
            off += x * (unsigned)L.stride;
            w *= wl[k * PG_WL + x];
          }
          e.w = w;
          e.off = (int)(off * 8u);  // byte offset
        }
        cols[idx] = e;
      }
    }
    __syncthreads();
    const int left = sd.N - c0;
    if (left > 0 && mode == 1) {
      // groups of 16 columns; this warp takes groups q, q + 4, ...; the next group's loads fly during the DMMAs
      const int ngr = ((((min(left, PG_CHUNK) + 3) >> 2) + KSTEP_GROUP - 1) / KSTEP_GROUP * KSTEP_GROUP) >> 2;
      const PgCol *cg = cols + side * PG_CHUNK + 4 * t;
      double v[MB][4], w[4];
      {
        const PgCol *e = cg + 16 * q;
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = e[j].w;
        const unsigned off = (unsigned)e[0].off;
#pragma unroll
        for (int I = 0; I < MB; ++I) pg_ldvec<4>(rowp[I] + off, v[I]);
      }
#pragma unroll 1
      for (int gr = q; gr < ngr; gr += PG_WPS) {
        double vn[MB][4], wn[4];
        {
          const PgCol *e = cg + 16 * (gr + PG_WPS < ngr ? gr + PG_WPS : gr);
#pragma unroll
          for (int j = 0; j < 4; ++j) wn[j] = e[j].w;
          const unsigned off = (unsigned)e[0].off;
#pragma unroll
          for (int I = 0; I < MB; ++I) pg_ldvec<4>(rowp[I] + off, vn[I]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int k = 0;
#pragma unroll
          for (int I = 0; I < MB; ++I) {
            const double a = v[I][j] * w[j];
#pragma unroll
            for (int J = 0; J <= I; ++J, ++k) dmma884(acc[k][0], acc[k][1], a, v[J][j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          w[j] = wn[j];
#pragma unroll
          for (int I = 0; I < MB; ++I) v[I][j] = vn[I][j];
        }
      }
    } else if (left > 0 && mode == 2) {
      if constexpr (MB == 2 || MB == 4) {
        const int nks_pad = (((min(left, PG_CHUNK) + 3) >> 2) + KSTEP_GROUP - 1) / KSTEP_GROUP * KSTEP_GROUP;
        const PgCol *cl = cols + side * PG_CHUNK + t;
        double v[PG_UNROLL][MB], w[PG_UNROLL];
#pragma unroll
        for (int u = 0; u < PG_UNROLL; ++u) {
          const PgCol e = cl[4 * (q + PG_WPS * u)];
          w[u] = e.w;
          pg_ldvec<MB>(rowp[0] + (unsigned)e.off, v[u]);
        }
#pragma unroll 1
        for (int ks0 = q; ks0 < nks_pad; ks0 += KSTEP_GROUP) {
          double vn[PG_UNROLL][MB], wn[PG_UNROLL];
          const int ksn = ks0 + KSTEP_GROUP < nks_pad ? ks0 + KSTEP_GROUP : ks0;
#pragma unroll
          for (int u = 0; u < PG_UNROLL; ++u) {
            const PgCol e = cl[4 * (ksn + PG_WPS * u)];
            wn[u] = e.w;
            pg_ldvec<MB>(rowp[0] + (unsigned)e.off, vn[u]);
          }
#pragma unroll
          for (int u = 0; u < PG_UNROLL; ++u) {
            int k = 0;
#pragma unroll
            for (int I = 0; I < MB; ++I) {
              const double a = v[u][I] * w[u];
#pragma unroll
              for (int J = 0; J <= I; ++J, ++k) dmma884(acc[k][0], acc[k][1], a, v[u][J]);
            }
          }
#pragma unroll
          for (int u = 0; u < PG_UNROLL; ++u) {
            w[u] = wn[u];
#pragma unroll
            for (int I = 0; I < MB; ++I) v[u][I] = vn[u][I];
          }
        }
      }
    } else if (left > 0) {
      const int nks_pad = (((min(left, PG_CHUNK) + 3) >> 2) + KSTEP_GROUP - 1) / KSTEP_GROUP * KSTEP_GROUP;
      const PgCol *cl = cols + side * PG_CHUNK + t;
      // software pipeline: the loads of trip n+1 are in flight while the DMMAs of trip n issue
      double v[PG_UNROLL][MB], w[PG_UNROLL];
#pragma unroll
      for (int u = 0; u < PG_UNROLL; ++u) {
        const PgCol e = cl[4 * (q + PG_WPS * u)];
        w[u] = e.w;
#pragma unroll
        for (int I = 0; I < MB; ++I) v[u][I] = __ldg((const double *)(rowp[I] + (unsigned)e.off));
      }
#pragma unroll 1
      for (int ks0 = q; ks0 < nks_pad; ks0 += KSTEP_GROUP) {
        double vn[PG_UNROLL][MB], wn[PG_UNROLL];
        const int ksn = ks0 + KSTEP_GROUP < nks_pad ? ks0 + KSTEP_GROUP : ks0;  // last trip reloads itself (L1 hit)
#pragma unroll
        for (int u = 0; u < PG_UNROLL; ++u) {
          const PgCol e = cl[4 * (ksn + PG_WPS * u)];
          wn[u] = e.w;
#pragma unroll
          for (int I = 0; I < MB; ++I) vn[u][I] = __ldg((const double *)(rowp[I] + (unsigned)e.off));
        }
#pragma unroll
        for (int u = 0; u < PG_UNROLL; ++u) {
          int k = 0;
#pragma unroll
          for (int I = 0; I < MB; ++I) {
            const double a = v[u][I] * w[u];
#pragma unroll
            for (int J = 0; J <= I; ++J, ++k) dmma884(acc[k][0], acc[k][1], a, v[u][J]);
          }
        }
#pragma unroll
        for (int u = 0; u < PG_UNROLL; ++u) {
          w[u] = wn[u];
#pragma unroll
          for (int I = 0; I < MB; ++I) v[u][I] = vn[u][I];
        }
      }
    }
  }
  // sum the four partial fragments of each side in warp order; tile (I,J) element [g][2t+{0,1}]
  int *modes = (int *)(red + 1);  // [2] fetch mode of each side (for the symmetric fill below)
  if (lane == 0 && q == 0) modes[side] = mode;
  double *gs = gram + side * MP * GLD;
  for (int wv = 0; wv < PG_WPS; ++wv) {
    __syncthreads();
    if (q == wv) {
      int k = 0;
#pragma unroll
      for (int I = 0; I < MB; ++I)
#pragma unroll
        for (int J = 0; J <= I; ++J, ++k) {
          // fragment (block I, row g) is bond-matrix row 8 I + g, or g MB + I in mode 2; same for the columns
          const int r = (mode == 2) ? g * MB + I : 8 * I + g;
          const int ca = (mode == 2) ? (2 * t) * MB + J : 8 * J + 2 * t;
          const int cb = (mode == 2) ? (2 * t + 1) * MB + J : 8 * J + 2 * t + 1;
          double *da = gs + r * GLD + ca, *db = gs + r * GLD + cb;
          double xa = 0.0, xb = 0.0;
          if (wv != 0) {
            xa = *da;
            xb = *db;
          }
          *da = xa + acc[k][0];
          *db = xb + acc[k][1];
        }
    }
  }
  __syncthreads();
  // strictly upper blocks by symmetry; G_j picks up lambda_e lambda_f on its bond indices
  for (int idx = tid; idx < 2 * MP * MP; idx += PG_THREADS) {
    const int sdx = idx / (MP * MP), rest = idx - sdx * MP * MP;
    const int r = rest / MP, c = rest - r * MP;
    const bool m2 = modes[sdx] == 2;
    const int br = m2 ? r % MB : r >> 3, bc = m2 ? c % MB : c >> 3;  // row-block ids: only blocks I >= J were computed
    if (br >= bc) {
      double *gp = gram + sdx * MP * GLD;
      double v = gp[r * GLD + c];
      if (sdx) v *= lamrow[r] * lamrow[c];
      gp[r * GLD + c] = v;
      if (br > bc) gp[c * GLD + r] = v;
    }
  }
  __syncthreads();

  // rho[i,j,i',j'] = sum_{e,f} G_i[(i,e),(i',f)] (lam_e lam_f G_j[(j,e),(j',f)]): one warp per entry, lanes over (e,f)
  const double *gi = gram, *gj = gram + MP * GLD;
  int eo[8];  // (e,f) pairs of this lane as offsets e*GLD + f; Dk <= 16 -> at most 8 per lane
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ef = lane + 32 * k;
    const int e = ef / Dk, f = ef - e * Dk;
    eo[k] = ef < Dk * Dk ? e * GLD + f : -1;
  }
  for (int idx = warp; idx < p.d4; idx += PG_THREADS / 32) {
    const int jp = idx % d, ip = (idx / d) % d, j = (idx / (d * d)) % d, i = idx / (d * d * d);
    const double *bi = gi + (i * Dk) * GLD + ip * Dk, *bj = gj + (j * Dk) * GLD + jp * Dk;
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (eo[k] >= 0) a = fma(bi[eo[k]], bj[eo[k]], a);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
    if (lane == 0) rho[idx] = a;
  }
  __syncthreads();
  if (warp == 0) {
    double tr = 0.0;
    for (int ij = lane; ij < d * d; ij += 32) tr += rho[ij * d * d + ij];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, m);
    if (lane == 0) red[0] = 1.0 / tr;
  }
  __syncthreads();
  const double itr = red[0];
  if (p.rdm_out != nullptr && ek == ek_first)
    for (int idx = tid; idx < p.d4; idx += PG_THREADS) p.rdm_out[(size_t)b * p.d4 + idx] = mkc(rho[idx] * itr, 0.0);
  if (p.pair_val != nullptr && warp == 0) {
    const double *op = p.ops + (size_t)b * p.op_point_stride + (size_t)ek * p.op_edge_stride;
    double v = 0.0;
    for (int idx = lane; idx < p.d4; idx += 32) v = fma(rho[idx] * itr, op[idx], v);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) p.pair_val[(size_t)b * p.m_edges + ek] = mkc(v, 0.0);
  }
}

// energy[b] = Re( sum_e pair_val[b][e] ) / n_tensors; cval[b] keeps the complex sum / n.
__global__ void pair_sum_kernel(const cplx *pair_val, int batch, int m_edges, int n_tensors, double *energy,
                                cplx *cval, const int *iter_state, const int *active) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  if (active != nullptr && active[b] == 0) return;
  if (iter_state != nullptr) {
    const int it = iter_state[b];
    if (!(it % 2 == 0 && it > 0)) return;
  }
  cplx s = mkc(0.0, 0.0);
  for (int e = 0; e < m_edges; ++e) s += pair_val[(size_t)b * m_edges + e];
  s = s * (1.0 / n_tensors);
  if (energy) energy[b] = s.x;
  if (cval) cval[b] = s;
}

// ---------------------------------------------------------------------------------------------
// one-site RDM: rho[s,s'] = sum_x T[s,x] conj(T[s',x]) prod_m lam_m[x_m]^2, trace-normalised.
// One CTA per (point, tensor).
// ---------------------------------------------------------------------------------------------
template <typename S>
__global__ void __launch_bounds__(RDM_THREADS) site_rdm_kernel(const SiteDesc *descs, int n_tensors, int d, int batch,
                                                               int m_edges, int dcap, const double *weights,
                                                               cplx *rdm_out /* [batch][n][d*d] */) {
  __shared__ double wleg[TNSU_MAX_LEGS * 64];
  __shared__ cplx red[RDM_THREADS / 32][TNSU_MAX_SPIN * TNSU_MAX_SPIN];
  __shared__ cplx tot[TNSU_MAX_SPIN * TNSU_MAX_SPIN];
  const int b = blockIdx.x / n_tensors;
  const int t = blockIdx.x - b * n_tensors;
  const SiteDesc &sd = descs[t];
  const int tid = threadIdx.x;
  const double *wpoint = weights + (size_t)b * m_edges * dcap;
  for (int idx = tid; idx < sd.nlegs * dcap; idx += RDM_THREADS) {
    int l = idx / dcap, x = idx - l * dcap;
    double w = (x < sd.dims[l]) ? wpoint[(size_t)sd.leg_edge[l] * dcap + x] : 1.0;
    wleg[l * 64 + x] = w * w;
  }
  __syncthreads();
  const S *tbase = (const S *)sd.base + (size_t)b * sd.point_stride;
  S acc[TNSU_MAX_SPIN][TNSU_MAX_SPIN];
#pragma unroll
  for (int i = 0; i < TNSU_MAX_SPIN; ++i)
#pragma unroll
    for (int j = 0; j < TNSU_MAX_SPIN; ++j) acc[i][j] = from_real<S>(0.0);
  for (long long x = tid; x < sd.nvirt; x += RDM_THREADS) {
    double w2 = 1.0;
    long long rem = x;
    for (int l = sd.nlegs - 1; l >= 0; --l) {
      const int dim = sd.dims[l];
      w2 *= wleg[l * 64 + (int)(rem % dim)];
      rem /= dim;
    }
    S v[TNSU_MAX_SPIN];
#pragma unroll
    for (int s = 0; s < TNSU_MAX_SPIN; ++s) v[s] = (s < d) ? tbase[(long long)s * sd.nvirt + x] : from_real<S>(0.0);
#pragma unroll
    for (int i = 0; i < TNSU_MAX_SPIN; ++i)
#pragma unroll
      for (int j = 0; j < TNSU_MAX_SPIN; ++j)
        if (i < d && j < d) fma_acc(acc[i][j], v[i] * w2, cj(v[j]));
  }
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int i = 0; i < TNSU_MAX_SPIN; ++i)
#pragma unroll
    for (int j = 0; j < TNSU_MAX_SPIN; ++j) {
      S a = acc[i][j];
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) a = a + shfl_xor(a, m);
      if (lane == 0) red[wid][i * TNSU_MAX_SPIN + j] = to_c128(a);
    }
  __syncthreads();
  if (tid < TNSU_MAX_SPIN * TNSU_MAX_SPIN) {
    cplx s = mkc(0.0, 0.0);
    for (int w = 0; w < RDM_THREADS / 32; ++w) s += red[w][tid];
    tot[tid] = s;
  }
  __syncthreads();
  if (tid < d * d) {
    cplx tr = mkc(0.0, 0.0);
    for (int s = 0; s < d; ++s) tr += tot[s * TNSU_MAX_SPIN + s];
    const double den = 1.0 / abs2(tr);
    const cplx itr = mkc(tr.x * den, -tr.y * den);
    const int i = tid / d, j = tid - i * d;
    rdm_out[((size_t)b * n_tensors + t) * d * d + tid] = tot[i * TNSU_MAX_SPIN + j] * itr;
  }
}

// T *= tscale (materialise the lazy normalisation before a download)
template <typename S>
__global__ void apply_scale_kernel(S *base, long long point_stride, long long elems, int batch, const double *tscale,
                                   int n_tensors, int t) {
  for (int b = blockIdx.y; b < batch; b += gridDim.y) {  // gridDim.y is capped at 65535: stride over the points
    const double sc = tscale[(size_t)b * n_tensors + t];
    S *ptr = base + (size_t)b * point_stride;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x)
      ptr[i] = ptr[i] * sc;
  }
}
__global__ void reset_scale_kernel(double *tscale, int n_tensors, int batch, int t) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) tscale[(size_t)b * n_tensors + t] = 1.0;
}
