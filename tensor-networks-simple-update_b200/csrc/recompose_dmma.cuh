// recompose_dmma.cuh — K3 of the real float64 path on the FP64 tensor cores.
//
// Replaces the recompose einsum P' = r~ . Q (reference src/tnsu/simple_update.py:274-275), the inverse-weight
// absorption (:290-291, tensor_network.py:264-278) and tensor_norm (:294-296, :487-496), like recompose_kernel.
//
//   P'[:, c] = [c < K] r~[:, c] - C . Y^H[:, c],   C = r~ S   (Mout x K, computed once per CTA)
//
// is a (Mout x K) x (K x N) product with Mout, K <= 32 and N in the hundreds: the small factor -C sits in registers
// as m8n8k4 A fragments, the reflector rows Y^H stream from HBM straight into B fragments (lane (g, t) loads
// Y^H[4 ks + t][c + g]: 64 contiguous bytes per row), the accumulators start at [c < K] r~ and leave through the
// epilogue (inverse environment weights, |.|^2 for the norm, store in the tensor's own layout).  A CTA is 4 warps;
// a warp trip covers 32 columns (4 tiles) with all its loads issued before the first DMMA.  Few registers -> many
// resident CTAs, so loads of one CTA overlap the stores of another: the kernel streams instead of alternating
// between a load phase and a store phase like recompose_kernel does.
#pragma once
#include "edge_update.cuh"

#define RCD_THREADS 128
#define RCD_TILES 4   // 8-column tiles per warp trip
#define RCD_WL 32     // leading dimension of the per-leg inverse-weight table

struct __align__(16) RcdCol {
  double inv_w;  // product of the inverse weights of the environment legs
  int out_off;   // element offset of (s = 0, x_k = 0) of this column in the updated tensor
  int pad;
};
struct __align__(16) RcdLeg {
  int dim;
  unsigned magic;  // 2^32 / dim rounded up: exact quotients for column indices below 2^20
  int stride;      // out_stride of the leg
  int pad;
};

DEV void rcd_dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

static inline int rcd_smem_bytes(int ncl, int ldm) {
  return (ncl + 8 * RCD_TILES) * 16 + TNSU_MAX_LEGS * 16 + TNSU_MAX_LEGS * RCD_WL * 8 + 32 * 4 + 3 * ldm * ldm * 8 + 64;
}

template <int MBO, int KB>
__global__ void __launch_bounds__(RCD_THREADS, (MBO * KB <= 4) ? 4 : 3) recompose_dmma_kernel(const EdgeLaunch<double> p,
                                                                                              int ncl_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int KS = 2 * KB;  // k-steps of 4
  const int CL = p.cluster;
  if (p.lq_flags != nullptr && p.lq_flags[blockIdx.x / CL] == 0) return;  // recomposed by recompose_chol_kernel (checked first: most CTAs end here)
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
  const int lt = threadIdx.x, lane = lt & 31, warp = lt >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int N = sd.N, K = sd.K, Mout = sd.Mout;
  const int Dnew = ed.Dnew;
  const int ldm = p.ldm;
  const int ncl = (N + CL - 1) / CL;
  const int c0 = rank * ncl;
  const int c_end = min(N, c0 + ncl);
  const int ncols = c_end - c0;

  RcdCol *cols = (RcdCol *)smem_raw;                        // [ncl_cap + 32]
  RcdLeg *legs = (RcdLeg *)(cols + ncl_cap + 8 * RCD_TILES);  // environment legs, innermost first
  double *winv = (double *)(legs + TNSU_MAX_LEGS);          // [legs][RCD_WL] inverse weights, same order
  int *rowout = (int *)(winv + TNSU_MAX_LEGS * RCD_WL);      // [32]
  double *matR = (double *)(rowout + 32);
  double *matS = matR + ldm * ldm;
  double *matC = matS + ldm * ldm;
  double *nred = matC + ldm * ldm;

  const double *wpoint = p.weights + (size_t)b * p.m_edges * p.dcap;
  // NB: lambda_k of THIS edge was just rewritten by K2; only the other legs' weights are used here
  for (int idx = lt; idx < TNSU_MAX_LEGS * RCD_WL; idx += RCD_THREADS) {
    const int k = idx / RCD_WL, x = idx - k * RCD_WL;
    int l = sd.nlegs - 1 - k;  // k-th environment leg from the inside: skip the bond leg
    if (l <= sd.bond_leg) --l;
    double w = 1.0;
    if (l >= 0) {
      const int dim = sd.dims[l];
      if (x < dim) w = 1.0 / wpoint[(size_t)sd.leg_edge[l] * p.dcap + x];  // np.power(w, -1), tensor_network.py:274
      if (x == 0) {
        RcdLeg L;
        L.dim = dim;
        L.magic = dim > 1 ? 0xFFFFFFFFu / (unsigned)dim + 1u : 0u;
        L.stride = (int)sd.out_stride[l];
        L.pad = 0;
        legs[k] = L;
      }
    }
    winv[idx] = w;
  }
  if (lt < 32) {
    const int s = lt / Dnew, x = lt - s * Dnew;
    rowout[lt] = (lt < Mout) ? (int)(s * sd.out_stride_phys + x * sd.out_stride[sd.bond_leg]) : 0;
  }
  const double *small = p.small + (size_t)item * SM_COUNT * ldm * ldm;
  const double *gR = small + (SM_R0 + side) * ldm * ldm;
  const double *gS = small + (SM_S0 + side) * ldm * ldm;
  for (int idx = lt; idx < Mout * ldm; idx += RCD_THREADS) matR[idx] = gR[idx];
  for (int idx = lt; idx < K * ldm; idx += RCD_THREADS) matS[idx] = gS[idx];
  __syncthreads();
  // column table of this slice, zero weight / offset 0 past its end
  const int nenv = sd.nlegs - 1;
  const int ntab = (ncols + 8 * RCD_TILES - 1) / (8 * RCD_TILES) * (8 * RCD_TILES);
  for (int cc = lt; cc < ntab; cc += RCD_THREADS) {
    RcdCol e;
    e.inv_w = 0.0;
    e.out_off = 0;
    e.pad = 0;
    if (cc < ncols) {
      unsigned rem = (unsigned)(c0 + cc);
      double w = 1.0;
      unsigned off = 0;
      for (int k = 0; k < nenv; ++k) {
        const RcdLeg L = legs[k];
        const unsigned qq = L.dim > 1 ? __umulhi(rem, L.magic) : rem;
        const unsigned x = rem - qq * (unsigned)L.dim;
        rem = qq;
        off += x * (unsigned)L.stride;
        w *= winv[k * RCD_WL + x];
      }
      e.inv_w = w;
      e.out_off = (int)off;
    }
    cols[cc] = e;
  }
  // C = r~ S  (Mout x K); S = Y1 T^H is lower triangular
  for (int idx = lt; idx < Mout * K; idx += RCD_THREADS) {
    const int i = idx / K, kp = idx - i * K;
    double a = 0.0;
    for (int c = kp; c < K; ++c) a = fma(matR[i * ldm + c], matS[c * ldm + kp], a);
    matC[i * ldm + kp] = a;
  }
  __syncthreads();

  // A fragments of -C: lane (g, t) holds -C[8 I + g][4 ks + t]
  double afr[MBO][KS];
#pragma unroll
  for (int I = 0; I < MBO; ++I)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int i = 8 * I + g, k = 4 * ks + t;
      afr[I][ks] = (i < Mout && k < K) ? -matC[i * ldm + k] : 0.0;
    }
  int ro[MBO];
#pragma unroll
  for (int I = 0; I < MBO; ++I) ro[I] = rowout[8 * I + g];
  // rows of Y^H this lane reads (clamped: A is zero there)
  const double *ybase = (const double *)sd.ybase + (size_t)b * sd.point_stride;
  const double *yrow[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int k = 4 * ks + t;
    yrow[ks] = ybase + (size_t)(k < K ? k : 0) * N;
  }
  double *tbase = (double *)sd.base + (size_t)b * sd.point_stride;
  double n2 = 0.0;
  for (int cg = warp * 8 * RCD_TILES; cg < ncols; cg += (RCD_THREADS / 32) * 8 * RCD_TILES) {
    double bfr[RCD_TILES][KS];
#pragma unroll
    for (int tl = 0; tl < RCD_TILES; ++tl) {
      const int c = min(c0 + cg + 8 * tl + g, c_end - 1);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) bfr[tl][ks] = __ldg(yrow[ks] + c);
    }
#pragma unroll
    for (int tl = 0; tl < RCD_TILES; ++tl) {
      const int cc = cg + 8 * tl + 2 * t;  // slice-local column of accumulator element 0 (element 1: cc + 1)
      const int ca = c0 + cc;
      const RcdCol e0 = cols[cc], e1 = cols[cc + 1];
#pragma unroll
      for (int I = 0; I < MBO; ++I) {
        const int i = 8 * I + g;
        double d0 = (i < Mout && ca < K) ? matR[i * ldm + ca] : 0.0;
        double d1 = (i < Mout && ca + 1 < K) ? matR[i * ldm + ca + 1] : 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) rcd_dmma(d0, d1, afr[I][ks], bfr[tl][ks]);
        if (i < Mout) {
          if (cc < ncols) {
            const double v = d0 * e0.inv_w;
            n2 = fma(v, v, n2);
            tbase[(long long)e0.out_off + ro[I]] = v;
          }
          if (cc + 1 < ncols) {
            const double v = d1 * e1.inv_w;
            n2 = fma(v, v, n2);
            tbase[(long long)e1.out_off + ro[I]] = v;
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, m);
  if (lane == 0) nred[warp] = n2;
  __syncthreads();
  if (lt == 0) {
    double tot = 0.0;
    for (int w = 0; w < RCD_THREADS / 32; ++w) tot += nred[w];
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

template <int MBO, int KB>
cudaError_t launch_recompose_dmma_v(const EdgeLaunch<double> &p, int n_items, int ncl_cap, cudaStream_t st) {
  auto kern = recompose_dmma_kernel<MBO, KB>;
  const int smem = rcd_smem_bytes(ncl_cap, p.ldm);
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  kern<<<2 * n_items * p.cluster, RCD_THREADS, smem, st>>>(p, ncl_cap);
  return cudaGetLastError();
}

// mout_max / k_max: largest Mout and K of the launch (both <= 32); ncl_cap: largest column count of a slice
inline cudaError_t launch_recompose_dmma(const EdgeLaunch<double> &p, int n_items, int mout_max, int k_max, int ncl_cap,
                                         cudaStream_t st) {
  const int mbo = (mout_max + 7) / 8, kb = (k_max + 7) / 8;
#define RCD_CASE(A, B) \
  if (mbo == A && kb == B) return launch_recompose_dmma_v<A, B>(p, n_items, ncl_cap, st);
  RCD_CASE(1, 1) RCD_CASE(1, 2) RCD_CASE(1, 3) RCD_CASE(1, 4)
  RCD_CASE(2, 1) RCD_CASE(2, 2) RCD_CASE(2, 3) RCD_CASE(2, 4)
  RCD_CASE(3, 1) RCD_CASE(3, 2) RCD_CASE(3, 3) RCD_CASE(3, 4)
  RCD_CASE(4, 1) RCD_CASE(4, 2) RCD_CASE(4, 3) RCD_CASE(4, 4)
#undef RCD_CASE
  return cudaErrorInvalidConfiguration;
}
