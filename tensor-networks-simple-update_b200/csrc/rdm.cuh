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
  int n_list;
  int batch, m_edges, dcap, d4;
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
  const int b = blockIdx.x / p.n_list;
  const int ek = p.edge_list[blockIdx.x - b * p.n_list];
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
struct SiteDesc {
  void *base;
  long long point_stride;
  int nlegs;
  int dims[TNSU_MAX_LEGS];
  int leg_edge[TNSU_MAX_LEGS];
  long long nvirt;  // prod dims
};

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
  const int b = blockIdx.y;
  const double sc = tscale[(size_t)b * n_tensors + t];
  S *ptr = base + (size_t)b * point_stride;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x)
    ptr[i] = ptr[i] * sc;
}
__global__ void reset_scale_kernel(double *tscale, int n_tensors, int batch, int t) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < batch) tscale[(size_t)b * n_tensors + t] = 1.0;
}
