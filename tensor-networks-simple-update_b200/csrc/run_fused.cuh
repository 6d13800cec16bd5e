// run_fused.cuh — SimpleUpdate.run() as ONE persistent kernel for networks that fit in shared memory.
//
// Replaces, for small networks (BASELINE configs 1 and 3: star D=2, "peps" cell D=4 — a whole network is 16 KiB), the
// launch chain of the general path: `for dt in dts: for i in range(max_iterations): simple_update(); every 2nd sweep
// compute_convergence_error() + energy_per_site()` (reference src/tnsu/simple_update.py:117-188).  One CTA owns one
// point (one network of the batch) from its current (dt, iteration) state to the end of its run:
//
//   state     all site tensors, all weight vectors, the lazy scales and the edge errors live in shared memory
//   edges     the dependency levels of the scheduler (engine.cu) are kept: the edges of one level share no tensor and
//             run concurrently, ONE WARP per edge, a block barrier between levels — the reference's sequential sweep
//             order is reproduced exactly as in the general path
//   one edge  absorb + stage the two bond matrices (dense, conflict-free) -> Householder LQ in shared memory (the
//             algorithm of lq_kernel: Gram row of the pivot by warp reduce-scatter, compact-WY factors) -> theta ->
//             the register-resident warp SVD (svd_warp.cuh, accumulating variant: this kernel is latency bound and
//             the accumulator's FMAs are free) -> recompose P' = [r~ 0] - (r~ S) Y^H, un-absorb, norm, back in place
//   checks    at every check iteration (i % 2 == 0 and i > 0) the warps share the m pair-RDM evaluations of
//             energy_per_site (:615-637), thread 0 runs the control flow of run() (:132-188) and logs
//
// No global-memory traffic and no launch between the first and the last sweep of a point.  Steady bond dimensions
// only (D_new == D_k on every edge): tnsu_run executes the bond-growth sweeps with the general kernels first.
#pragma once
#include "edge_update.cuh"

#define FUSED_MAX_WARPS 4

// per-point control state of the batched run loop (one entry per point; shared with run_state_kernel in engine.cu)
struct RunState {
  int batch, m_edges, n_dts, max_iter, log_cap;
  double tol;
  int *slot, *iter, *step, *active, *conv, *nchecks, *log_slot, *log_step, *nactive;
  const int *is_last;
  double *log_err, *log_energy;
  const double *edge_err, *energy;
};

struct FusedLaunch {
  const EdgeDesc *descs;    // steady-state descriptors, indexed by edge id
  const SiteDesc *sites;    // per tensor: global base, point stride, element count
  const int *level_edges;   // edge ids, level-major
  const int *level_off;     // [n_levels + 1]
  const int *toff;          // [n_tensors] offset (doubles) of tensor t in the CTA's shared-memory state
  int n_levels, n_warps;
  int batch, m_edges, n_tensors, dcap, d4, d;
  double *weights;          // [batch][m][dcap]
  double *tscale;           // [batch][n]
  double *edge_err;         // [batch][m]
  double *energy;           // [batch]
  const double *gates;      // [slot][batch][m][d4]
  long long gate_slot_stride;
  const double *hams;       // [batch][m][d4]
  RunState rs;
  int *status;
  long long *prof;          // optional [8] cycle counters of point 0 / warp 0 (TNSU_FUSED_PROFILE=1): stage, LQ, SVD, recompose, energy, level barriers, control
  // geometry of the shared-memory layout (doubles)
  int tens_doubles;         // all tensors
  int warp_doubles;         // per-warp work area
  int ldm, ldp, mrows;      // small-matrix leading dimension; staged bond matrix: leading dimension and rows (multiple of 8)
  int svd_np, svd_precond, svd_ldt, svd_group_doubles, svd_max_sweeps;
  int svd_accv, svd_tb_off;  // accumulate the right rotations (no preconditioner: tiny thetas) or V-free with the transpose buffer at tb_off
  int descs_in_smem;         // the edge descriptors are copied into shared memory (m * sizeof(EdgeDesc) bytes behind the warp areas)
};

// per-warp work area
struct FusedWarp {
  double *gs;                 // SVD group memory: L_i | L_j | gate | lam_old | theta / transpose buffer
  double *S[2], *Rt[2];       // compact-WY factor S = Y1 T^H and r~ per side, [.][ldm]
  double *Z, *T;              // LQ scratch (K x K each); Z is reused for C = r~ S in the recompose
  double *Pd[2];              // staged bond matrices / reflectors Y^H, [mrows][ldp]
  double *wrow, *tau;         // [32] each
  int *rowoff;                // [32]
};

DEV FusedWarp fused_warp_mem(const FusedLaunch &p, double *base) {
  FusedWarp w;
  const int ldm2 = p.ldm * p.ldm;
  w.gs = base;
  double *q = base + p.svd_group_doubles;
  w.S[0] = q;
  w.S[1] = q + ldm2;
  w.Rt[0] = q + 2 * ldm2;
  w.Rt[1] = q + 3 * ldm2;
  w.Z = q + 4 * ldm2;
  w.T = q + 5 * ldm2;
  q += 6 * ldm2;
  w.Pd[0] = q;
  w.Pd[1] = q + (size_t)p.mrows * p.ldp;
  q += 2 * (size_t)p.mrows * p.ldp;
  w.wrow = q;
  w.tau = q + 32;
  w.rowoff = (int *)(q + 64);
  return w;
}
static inline int fused_warp_doubles(int svd_group_doubles, int ldm, int mrows, int ldp) {
  return ((svd_group_doubles + 6 * ldm * ldm + 2 * mrows * ldp + 64 + 16) + 1) & ~1;
}

// run_fused_kernel lives in its own translation unit (fused_inst.cu, which defines RUN_FUSED_IMPL); engine.cu only needs
// the launch structures above and this entry point
// blocks_per_sm (optional) receives the kernel variant's occupancy; with dry_run nothing is launched
cudaError_t launch_run_fused_dispatch(const FusedLaunch &p, int svd_rows, int max_m, int max_n, int smem_bytes, cudaStream_t st,
                                      bool dry_run, int *blocks_per_sm);

#ifdef RUN_FUSED_IMPL
#include "svd_warp.cuh"

// element offset of column c (s = 0, x_k = 0) and the product of the environment weights on the other legs
// `wts` is either the weight table (absorb, tensor_network.py:242-259) or the table of their reciprocals 1 / lambda
// (un-absorb, np.power(w, -1) leg by leg, :261-278), which the kernel keeps beside the weights
DEV void fused_decode(const SideDesc &sd, int c, const double *__restrict__ wts, int dcap, int &off, double &w) {
  off = 0;
  w = 1.0;
  unsigned rem = (unsigned)c;
  for (int l = sd.nlegs - 1; l >= 0; --l) {
    if (l == sd.bond_leg) continue;
    unsigned x;
    divmod_small(rem, (unsigned)sd.dims[l], rem, x);
    off += (int)x * (int)sd.in_stride[l];
    w *= wts[sd.leg_edge[l] * dcap + x];
  }
}

// Pd[r][c] = T[s, x_k, others] * prod_{m != k} lambda_m[x_m] * scale   (r = s * D_k + x_k); rows M..mrows-1 are zero
DEV void fused_stage(const FusedLaunch &p, const SideDesc &sd, int Dk, const double *tens_t, const double *wts, double scale,
                     double *Pd, int *rowoff, int lane) {
  const int M = sd.M, N = sd.N;
  if (lane < p.mrows && lane < 32) {
    const int r = lane < M ? lane : 0;
    const int s = r / Dk, x = r - s * Dk;
    rowoff[lane] = (int)(s * sd.in_stride_phys + x * sd.in_stride[sd.bond_leg]);
  }
  __syncwarp();
  for (int c = lane; c < N; c += 32) {
    int off;
    double w;
    fused_decode(sd, c, wts, p.dcap, off, w);
    w *= scale;
    const double *src = tens_t + off;
    for (int r = 0; r < M; ++r) Pd[(size_t)r * p.ldp + c] = src[rowoff[r]] * w;
    for (int r = M; r < p.mrows; ++r) Pd[(size_t)r * p.ldp + c] = 0.0;
  }
  __syncwarp();
}

// Householder LQ of the staged bond matrix (rows reduced one by one, the algorithm of lq_kernel / lq_rolled_kernel):
// P = L Q.  On return Pd holds the reflectors Y^H (row k: zeros | head | tail), L (M x K) and S = Y1 T^H (K x K) are
// written with leading dimension ldm.
DEV void fused_lq(const FusedLaunch &p, const FusedWarp &wm, double *Pd, int M, int N, int K, double *L, double *S, int lane) {
  const int ldp = p.ldp, ldm = p.ldm;
  double *Z = wm.Z, *T = wm.T;
  for (int idx = lane; idx < M * ldm; idx += 32) L[idx] = 0.0;
  for (int idx = lane; idx < K * ldm; idx += 32) {
    Z[idx] = 0.0;
    T[idx] = 0.0;
  }
  __syncwarp();
  for (int kk = 0; kk < K; ++kk) {
    // Gram row of the pivot with every row, over my columns c > kk; lane i ends up with the value of row i
    double g = 0.0;
    const double *prow = Pd + (size_t)kk * ldp;
    for (int cb = 0; cb < p.mrows; cb += 8) {
      double acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.0;
      for (int c = lane; c < N; c += 32) {
        if (c > kk) {
          const double pk = prow[c];
          const double *col = Pd + (size_t)cb * ldp + c;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fma(col[(size_t)i * ldp], pk, acc[i]);
        }
      }
      const double tot = warp_reduce_scatter<double, 8>(acc, lane);
      if ((lane >> 3) == (cb >> 3)) g = tot;
    }
    const double mycol = (lane < M) ? Pd[(size_t)lane * ldp + kk] : 0.0;
    const double alpha = prow[kk];
    const double gk = __shfl_sync(0xffffffffu, g, kk);
    const double norm2 = fma(alpha, alpha, gk);
    double w_l = 0.0, t_l = g, lnew = mycol, head = alpha, lkk = alpha, tau = 0.0;
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
    if (lane < M) {
      if (lane > kk)
        L[lane * ldm + kk] = lnew;
      else if (lane == kk) {
        L[kk * ldm + kk] = lkk;
        wm.tau[kk] = tau;
      } else
        Z[lane * ldm + kk] = t_l;
    }
    wm.wrow[lane] = (lane > kk && lane < M) ? w_l : 0.0;
    __syncwarp();
    // rank-1 update of the rows still to be reduced
    for (int c = lane; c < N; c += 32) {
      if (c > kk) {
        const double pk = prow[c];
        for (int i = kk + 1; i < M; ++i) {
          double *q = Pd + (size_t)i * ldp + c;
          *q = fma(-wm.wrow[i], pk, *q);
        }
      }
    }
    if (lane == 0) Pd[(size_t)kk * ldp + kk] = head;
    __syncwarp();
  }
  // compact-WY factor: T[:k,k] = -tau_k T[:k,:k] Z[:k,k], T[k,k] = tau_k   (lane r owns row r of T)
  for (int kk = 0; kk < K; ++kk) {
    const double tk = wm.tau[kk];
    if (lane < kk) {
      double a0 = 0.0, a1 = 0.0;
      int j = 0;
      for (; j + 1 < kk; j += 2) {
        a0 = fma(T[lane * ldm + j], Z[j * ldm + kk], a0);
        a1 = fma(T[lane * ldm + j + 1], Z[(j + 1) * ldm + kk], a1);
      }
      if (j < kk) a0 = fma(T[lane * ldm + j], Z[j * ldm + kk], a0);
      T[lane * ldm + kk] = -((a0 + a1) * tk);
    } else if (lane == kk) {
      T[kk * ldm + kk] = tk;
    }
  }
  __syncwarp();
  // S[c][k'] = sum_{k = k'..c} Y^H[k][c] T[k'][k]   (the first K columns of Y^H; entries left of the head are zero)
  for (int idx = lane; idx < K * K; idx += 32) {
    const int c = idx / K, kp = idx - c * K;
    double a = 0.0;
    for (int k = kp; k <= c; ++k) a = fma(Pd[(size_t)k * ldp + c], T[kp * ldm + k], a);
    S[c * ldm + kp] = a;
  }
  __syncwarp();
}

// P'[:, c] = [c < K] r~[:, c] - (r~ S) Y^H[:, c], un-absorbed, stored in place; returns 1 / ||T'||_F (the new lazy scale)
DEV double fused_recompose(const FusedLaunch &p, const FusedWarp &wm, const SideDesc &sd, const double *Pd, const double *Rt,
                           const double *S, double *tens_t, const double *iwts, int lane) {
  const int ldp = p.ldp, ldm = p.ldm;
  const int N = sd.N, K = sd.K, Mout = sd.Mout;
  double *C = wm.Z;
  for (int idx = lane; idx < Mout * K; idx += 32) {
    const int i = idx / K, kp = idx - i * K;
    double a = 0.0;
    for (int c = kp; c < K; ++c) a = fma(Rt[i * ldm + c], S[c * ldm + kp], a);  // S is lower triangular
    C[i * ldm + kp] = a;
  }
  __syncwarp();
  double n2 = 0.0;
  for (int c = lane; c < N; c += 32) {
    int off;
    double iw;
    fused_decode(sd, c, iwts, p.dcap, off, iw);
    double *dst = tens_t + off;
    const int kmax = (c < K - 1) ? c : K - 1;  // Y^H[k][c] = 0 for k > c
    for (int i = 0; i < Mout; ++i) {
      double o0 = (c < K) ? Rt[i * ldm + c] : 0.0, o1 = 0.0;
      const double *ci = C + i * ldm;
      int k = 0;
      for (; k + 1 <= kmax; k += 2) {
        o0 = fma(-ci[k], Pd[(size_t)k * ldp + c], o0);
        o1 = fma(-ci[k + 1], Pd[(size_t)(k + 1) * ldp + c], o1);
      }
      if (k <= kmax) o0 = fma(-ci[k], Pd[(size_t)k * ldp + c], o0);
      const double v = (o0 + o1) * iw;
      n2 = fma(v, v, n2);
      dst[wm.rowoff[i]] = v;
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, m);
  __syncwarp();
  return 1.0 / sqrt(n2);
}


// ---- register-resident variants for narrow bond matrices (N <= 32 * CPL columns, M <= MMAX rows): every lane owns CPL whole
// columns of P in registers from the absorbing load to the finished reflectors (the algorithm of lq_kernel with one warp) ----
template <int MMAX, int CPL>
DEV void fused_lq_regs(const FusedLaunch &p, const FusedWarp &wm, const SideDesc &sd, int Dk, const double *__restrict__ tens_t,
                       const double *__restrict__ wts, double scale, double *__restrict__ Pd, double *__restrict__ L,
                       double *__restrict__ S, int lane) {
  const int ldp = p.ldp, ldm = p.ldm;
  const int M = sd.M, N = sd.N, K = sd.K;
  double *__restrict__ Z = wm.Z, *__restrict__ T = wm.T;
  double *colk = wm.wrow;  // [32] the pivot column of the current step
  if (lane < MMAX) {
    const int r = lane < M ? lane : 0;
    const int s = r / Dk, x = r - s * Dk;
    wm.rowoff[lane] = (int)(s * sd.in_stride_phys + x * sd.in_stride[sd.bond_leg]);
  }
  for (int idx = lane; idx < M * ldm; idx += 32) L[idx] = 0.0;
  for (int idx = lane; idx < K * ldm; idx += 32) {
    Z[idx] = 0.0;
    T[idx] = 0.0;
  }
  __syncwarp();
  double col[CPL][MMAX];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    int off;
    double w;
    fused_decode(sd, c < N ? c : 0, wts, p.dcap, off, w);
    w = (c < N) ? w * scale : 0.0;
    const double *src = tens_t + off;
#pragma unroll
    for (int r = 0; r < MMAX; ++r) col[j][r] = (r < M) ? src[wm.rowoff[r]] * w : 0.0;
  }
  const int q = lane & (MMAX - 1);  // the row this lane finalises (lanes >= MMAX duplicate)
  static_for<0, MMAX>([&](auto kk_c) {
    constexpr int kk = decltype(kk_c)::value;
    if (kk < K) {
      if (lane == kk) {  // column kk is column j = 0 of lane kk (kk < 32)
#pragma unroll
        for (int i = 0; i < MMAX; ++i) colk[i] = col[0][i];
      }
      double acc[MMAX];
#pragma unroll
      for (int i = 0; i < MMAX; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const double pk = (lane + 32 * j > kk) ? col[j][kk] : 0.0;
          a = fma(col[j][i], pk, a);
        }
        acc[i] = a;
      }
      const double g = warp_reduce_scatter<double, MMAX>(acc, lane);  // lane l: total of row l % MMAX
      __syncwarp();
      const double mycol = colk[q];
      const double alpha = colk[kk];
      const double gk = __shfl_sync(0xffffffffu, g, kk);
      const double norm2 = fma(alpha, alpha, gk);
      double w_l = 0.0, t_l = g, lnew = mycol, head = alpha, lkk = alpha, tau = 0.0;
      if (norm2 > 0.0) {
        const double norm = norm2 * rsqrt(norm2);
        const double absa = fabs(alpha);
        const double phase = (alpha < 0.0) ? -1.0 : 1.0;
        const double uk = fma(phase, norm, alpha);
        tau = 1.0 / fma(norm, absa, norm2);
        t_l = fma(mycol, uk, g);
        w_l = t_l * tau;
        lnew = fma(-w_l, uk, mycol);
        head = uk;
        lkk = -(phase * norm);
      }
      if (lane < MMAX && lane < M) {
        if (lane > kk)
          L[lane * ldm + kk] = lnew;
        else if (lane == kk) {
          L[kk * ldm + kk] = lkk;
          wm.tau[kk] = tau;
        } else
          Z[lane * ldm + kk] = t_l;
      }
      if (!(q > kk && q < M)) w_l = 0.0;
#pragma unroll
      for (int i = kk + 1; i < MMAX; ++i) {
        const double wi = __shfl_sync(0xffffffffu, w_l, i);
#pragma unroll
        for (int j = 0; j < CPL; ++j)
          if (lane + 32 * j > kk) col[j][i] = fma(-wi, col[j][kk], col[j][i]);
      }
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        if (c == kk)
          col[j][kk] = head;
        else if (c < kk)
          col[j][kk] = 0.0;
      }
      __syncwarp();  // colk is rewritten by the next step
    }
  });
  // the reflectors Y^H (K x N) wait in shared memory for the recompose
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    if (c < N) {
#pragma unroll
      for (int k = 0; k < MMAX; ++k)
        if (k < K) Pd[(size_t)k * ldp + c] = col[j][k];
    }
  }
  __syncwarp();
  for (int kk = 0; kk < K; ++kk) {
    const double tk = wm.tau[kk];
    if (lane < kk) {
      double a0 = 0.0, a1 = 0.0;
      int j = 0;
      for (; j + 1 < kk; j += 2) {
        a0 = fma(T[lane * ldm + j], Z[j * ldm + kk], a0);
        a1 = fma(T[lane * ldm + j + 1], Z[(j + 1) * ldm + kk], a1);
      }
      if (j < kk) a0 = fma(T[lane * ldm + j], Z[j * ldm + kk], a0);
      T[lane * ldm + kk] = -((a0 + a1) * tk);
    } else if (lane == kk) {
      T[kk * ldm + kk] = tk;
    }
  }
  __syncwarp();
  for (int idx = lane; idx < K * K; idx += 32) {
    const int c = idx / K, kp = idx - c * K;
    double a = 0.0;
    for (int k = kp; k <= c; ++k) a = fma(Pd[(size_t)k * ldp + c], T[kp * ldm + k], a);
    S[c * ldm + kp] = a;
  }
  __syncwarp();
}

template <int MMAX, int CPL>
DEV double fused_recompose_regs(const FusedLaunch &p, const FusedWarp &wm, const SideDesc &sd, const double *__restrict__ Pd,
                                const double *__restrict__ Rt, const double *__restrict__ S, double *__restrict__ tens_t,
                                const double *__restrict__ iwts, int lane) {
  const int ldp = p.ldp, ldm = p.ldm;
  const int N = sd.N, K = sd.K, Mout = sd.Mout;
  double *__restrict__ C = wm.Z;
  for (int idx = lane; idx < Mout * K; idx += 32) {
    const int i = idx / K, kp = idx - i * K;
    double a = 0.0;
    for (int c = kp; c < K; ++c) a = fma(Rt[i * ldm + c], S[c * ldm + kp], a);
    C[i * ldm + kp] = a;
  }
  __syncwarp();
  double n2 = 0.0;
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    if (c < N) {
      int off;
      double iw;
      fused_decode(sd, c, iwts, p.dcap, off, iw);
      double yk[MMAX];
#pragma unroll
      for (int k = 0; k < MMAX; ++k) yk[k] = (k < K && k <= c) ? Pd[(size_t)k * ldp + c] : 0.0;
      double *dst = tens_t + off;
      for (int i = 0; i < Mout; ++i) {
        const double *ci = C + i * ldm;
        double o0 = (c < K) ? Rt[i * ldm + c] : 0.0, o1 = 0.0;
#pragma unroll
        for (int k = 0; k < MMAX; k += 2) {
          if (k < K) o0 = fma(-ci[k], yk[k], o0);
          if (k + 1 < K) o1 = fma(-ci[k + 1], yk[k + 1], o1);
        }
        const double v = (o0 + o1) * iw;
        n2 = fma(v, v, n2);
        dst[wm.rowoff[i]] = v;
      }
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, m);
  __syncwarp();
  return 1.0 / sqrt(n2);
}

// one imaginary-time-evolution step on edge ek (reference :219-296), executed by one warp (or shared by two)
// `wm` is the work area that holds the edge's data (L, S, r~, the staged bond matrices, the SVD group memory); `own` is the
// calling warp's scratch (Z / T / pivot column / row offsets).  role 0: this warp does the whole update (wm == own).
// roles 1 / 2: the level has ONE edge and two warps share it — warp 0 (role 1) takes side 0 and the SVD, warp 1 (role 2)
// takes side 1; the three phases are separated by block barriers that every warp of the CTA executes.
template <int R, int MMAX, int CPL, bool ACCV>
DEV void fused_edge_update(const FusedLaunch &p, const EdgeDesc *descs, const FusedWarp &wm, const FusedWarp &own, int role,
                           double *tens, double *wts, double *iwts, double *tsc, double *eerr, int b, int ek, int slot, int lane) {
  const EdgeDesc &ed = descs[ek];
  const int ldm2 = p.ldm * p.ldm;
  double *L[2] = {wm.gs, wm.gs + ldm2};
  const bool prof = p.prof != nullptr && b == 0 && (threadIdx.x >> 5) == 0 && lane == 0;
  const bool mine[2] = {role != 2, role != 1};  // the sides this warp reduces and recomposes
  long long c0 = clock64();
  for (int side = 0; side < 2; ++side) {
    if (!mine[side]) continue;
    const SideDesc &sd = ed.s[side];
    long long c1;
    if constexpr (MMAX > 0) {
      c1 = c0;
      fused_lq_regs<MMAX, CPL>(p, own, sd, ed.Dk, tens + p.toff[sd.tensor], wts, tsc[sd.tensor], wm.Pd[side], L[side], wm.S[side], lane);
    } else {
      fused_stage(p, sd, ed.Dk, tens + p.toff[sd.tensor], wts, tsc[sd.tensor], wm.Pd[side], own.rowoff, lane);
      c1 = clock64();
      fused_lq(p, own, wm.Pd[side], sd.M, sd.N, sd.K, L[side], wm.S[side], lane);
    }
    const long long c2 = clock64();
    if (prof) {
      p.prof[0] += c1 - c0;
      p.prof[1] += c2 - c1;
    }
    c0 = c2;
  }
  if (role != 0) bar_sync(1, blockDim.x);  // both L factors are in the edge's work area
  if (role != 2) {
  const double *gate_src = p.gates + (size_t)slot * p.gate_slot_stride + ((size_t)b * p.m_edges + ek) * p.d4;
  int theta_ready = 0;
  if (p.svd_precond) {
    // theta^T (the preconditioned SVD always iterates A' = theta^T) by all 32 lanes; svdw_body would build it with the
    // 2 * NP lanes of its group only.  theta[(qi,i'),(j',qj)] = sum_{i,j,e} L_i[(i,e),qi] lam_e L_j[(j,e),qj] G[i,j,i',j']
    const int d = ed.d, Dk = ed.Dk, Ki = ed.s[0].K, Kj = ed.s[1].K, ldm = p.ldm, lda = 2 * R;
    double *gate_s = wm.gs + 2 * ldm2, *lam_old = gate_s + p.d4, *th = lam_old + p.dcap;
    const double *Li = L[0], *Lj = L[1];
    for (int idx = lane; idx < p.d4; idx += 32) gate_s[idx] = gate_src[idx];
    for (int x = lane; x < Dk; x += 32) lam_old[x] = wts[ek * p.dcap + x];
    if constexpr (!ACCV)
      for (int idx = lane; idx < lda * 2 * R; idx += 32) th[idx] = 0.0;
    __syncwarp();
    for (int idx = lane; idx < Ki * Kj; idx += 32) {
      const int qi = idx / Kj, qj = idx - qi * Kj;
      double out[TNSU_MAX_SPIN][TNSU_MAX_SPIN];
#pragma unroll
      for (int ip = 0; ip < TNSU_MAX_SPIN; ++ip)
#pragma unroll
        for (int jp = 0; jp < TNSU_MAX_SPIN; ++jp) out[ip][jp] = 0.0;
#pragma unroll 1
      for (int ij = 0; ij < d * d; ++ij) {
        const int i = ij / d, j = ij - i * d;
        double acc = 0.0;
        for (int e = 0; e < Dk; ++e) acc = fma(Li[(i * Dk + e) * ldm + qi] * lam_old[e], Lj[(j * Dk + e) * ldm + qj], acc);
        const double *gp = gate_s + ij * d * d;
#pragma unroll
        for (int ip = 0; ip < TNSU_MAX_SPIN; ++ip)
#pragma unroll
          for (int jp = 0; jp < TNSU_MAX_SPIN; ++jp)
            if (ip < d && jp < d) out[ip][jp] = fma(acc, gp[ip * d + jp], out[ip][jp]);
      }
#pragma unroll
      for (int ip = 0; ip < TNSU_MAX_SPIN; ++ip)
#pragma unroll
        for (int jp = 0; jp < TNSU_MAX_SPIN; ++jp)
          if (ip < d && jp < d) th[(size_t)(qi * d + ip) * lda + jp * Kj + qj] = out[ip][jp];
    }
    __syncwarp();
    theta_ready = 1;
  }
  SvdwArgs a{};
  a.theta_ready = theta_ready;
  a.NP = p.svd_np;
  a.n_groups = 1;
  a.valid = lane < 2 * p.svd_np;
  a.d = ed.d;
  a.Dk = ed.Dk;
  a.Dnew = ed.Dnew;
  a.Ki = ed.s[0].K;
  a.Kj = ed.s[1].K;
  a.ni = ed.ni;
  a.nj = ed.nj;
  a.Mi = ed.s[0].M;
  a.Mj = ed.s[1].M;
  a.gs = wm.gs;
  a.ldm = p.ldm;
  a.dcap = p.dcap;
  a.d4 = p.d4;
  a.precond = p.svd_precond;
  a.ldt = p.svd_ldt;
  a.tb_off = p.svd_tb_off;
  a.max_sweeps = p.svd_max_sweeps;
  a.src_li = L[0];
  a.src_lj = L[1];
  a.gate_src = gate_src;
  a.lam = wts + ek * p.dcap;
  a.edge_err = eerr + ek;
  a.rt_i = wm.Rt[0];
  a.rt_j = wm.Rt[1];
  a.status = p.status;
  if (prof) {  // phase cycles of this warp's SVD: dbg_info[8..10] = theta + QR | Jacobi | extraction
    a.dbg_info = (int *)(p.prof + 8);
    a.dbg_sigma = (double *)(p.prof + 16);
  }
  svdw_body<R, ACCV>(a);
  __syncwarp();
  if (lane < ed.Dnew) iwts[ek * p.dcap + lane] = 1.0 / wts[ek * p.dcap + lane];
  __syncwarp();
  {
    const long long c1 = clock64();
    if (prof) p.prof[2] += c1 - c0;
    c0 = c1;
  }
  }  // role != 2
  if (role != 0) bar_sync(1, blockDim.x);  // r~_i, r~_j, lambda' and its reciprocals are complete
  for (int side = 0; side < 2; ++side) {
    if (!mine[side]) continue;
    const SideDesc &sd = ed.s[side];
    // row offsets of this side (the staging of side 1 overwrote those of side 0)
    if (lane < 32) {
      const int r = lane < sd.Mout ? lane : 0;
      const int s = r / ed.Dnew, x = r - s * ed.Dnew;
      own.rowoff[lane] = (int)(s * sd.out_stride_phys + x * sd.out_stride[sd.bond_leg]);
    }
    __syncwarp();
    double sc;
    if constexpr (MMAX > 0)
      sc = fused_recompose_regs<MMAX, CPL>(p, own, sd, wm.Pd[side], wm.Rt[side], wm.S[side], tens + p.toff[sd.tensor], iwts, lane);
    else
      sc = fused_recompose(p, own, sd, wm.Pd[side], wm.Rt[side], wm.S[side], tens + p.toff[sd.tensor], iwts, lane);
    if (lane == 0) tsc[sd.tensor] = sc;
    __syncwarp();
  }
  if (prof) p.prof[3] += clock64() - c0;
}

// <H_e> on edge ek from the two-site RDM (reference :525-593, :605-613), executed by one warp; the value is the real part
// of sum rho * H with rho trace-normalised
DEV double fused_pair_expectation(const FusedLaunch &p, const EdgeDesc *descs, const FusedWarp &wm, const double *tens, const double *wts,
                                  const double *tsc, int b, int ek, int lane) {
  const EdgeDesc &ed = descs[ek];
  const int d = ed.d, Dk = ed.Dk, M = d * Dk;
  double *gram[2] = {wm.S[0], wm.Rt[0]};  // S[0], S[1] and Rt[0], Rt[1] are adjacent: 2 * ldm^2 >= M^2 doubles each
  for (int side = 0; side < 2; ++side) {
    const SideDesc &sd = ed.s[side];
    fused_stage(p, sd, Dk, tens + p.toff[sd.tensor], wts, tsc[sd.tensor], wm.Pd[side], wm.rowoff, lane);
    const double *Pd = wm.Pd[side];
    for (int idx = lane; idx < M * M; idx += 32) {
      const int r = idx / M, q = idx - r * M;
      double a0 = 0.0, a1 = 0.0;
      const double *pr = Pd + (size_t)r * p.ldp, *pq = Pd + (size_t)q * p.ldp;
      int c = 0;
      for (; c + 1 < sd.N; c += 2) {
        a0 = fma(pr[c], pq[c], a0);
        a1 = fma(pr[c + 1], pq[c + 1], a1);
      }
      if (c < sd.N) a0 = fma(pr[c], pq[c], a0);
      gram[side][idx] = a0 + a1;
    }
  }
  __syncwarp();
  // rho[i,j,i',j'] = sum_{e,e'} lam_e lam_e' G_i[(i,e),(i',e')] G_j[(j,e),(j',e')]
  const double *lam = wts + ek * p.dcap;
  const double *gi = gram[0], *gj = gram[1];
  const double *op = p.hams + ((size_t)b * p.m_edges + ek) * p.d4;
  double tr = 0.0, val = 0.0;
  for (int idx = lane; idx < p.d4; idx += 32) {
    const int jp = idx % d, ip = (idx / d) % d, j = (idx / (d * d)) % d, i = idx / (d * d * d);
    double a = 0.0;
    for (int e = 0; e < Dk; ++e)
      for (int f = 0; f < Dk; ++f) a = fma(gi[(i * Dk + e) * M + ip * Dk + f] * (lam[e] * lam[f]), gj[(j * Dk + e) * M + jp * Dk + f], a);
    if (i == ip && j == jp) tr += a;
    val = fma(a, op[idx], val);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    tr += __shfl_xor_sync(0xffffffffu, tr, m);
    val += __shfl_xor_sync(0xffffffffu, val, m);
  }
  __syncwarp();
  return val / tr;
}

template <int R, int MMAX, int CPL, bool ACCV>
__global__ void __launch_bounds__(32 * FUSED_MAX_WARPS) run_fused_kernel(const FusedLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  if (p.rs.active[b] == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  const int m = p.m_edges, n = p.n_tensors;

  // shared memory: tensors | weights [m][dcap] | 1 / weights [m][dcap] | tscale [n] | edge_err [m] | pair values [m] |
  // control [8] | warp areas | edge descriptors
  double *tens = (double *)smem_raw;
  double *wts = tens + p.tens_doubles;
  double *iwts = wts + m * p.dcap;
  double *tsc = iwts + m * p.dcap;
  double *eerr = tsc + n;
  double *pval = eerr + m;
  int *ctrl = (int *)(pval + m);
  double *warp_base = pval + m + 8;
  const FusedWarp wm = fused_warp_mem(p, warp_base + (size_t)warp * p.warp_doubles);
  const FusedWarp wm0 = fused_warp_mem(p, warp_base);  // warp 0's area: the edge data of a level whose single edge two warps share
  const EdgeDesc *descs = p.descs;
  if (p.descs_in_smem) {
    unsigned long long *dst = (unsigned long long *)(warp_base + (size_t)p.n_warps * p.warp_doubles);
    const unsigned long long *src = (const unsigned long long *)p.descs;
    for (int idx = tid; idx < (int)(m * sizeof(EdgeDesc) / 8); idx += nthr) dst[idx] = src[idx];
    descs = (const EdgeDesc *)dst;
  }

  for (int t = 0; t < n; ++t) {
    const SiteDesc &sd = p.sites[t];
    const double *src = (const double *)sd.base + (size_t)b * sd.point_stride;
    const int elems = (int)sd.nvirt * p.d;
    double *dst = tens + p.toff[t];
    for (int idx = tid; idx < elems; idx += nthr) dst[idx] = src[idx];
  }
  for (int idx = tid; idx < m * p.dcap; idx += nthr) {
    const double w = p.weights[(size_t)b * m * p.dcap + idx];
    wts[idx] = w;
    iwts[idx] = 1.0 / w;  // entries beyond an edge's dimension are never read
  }
  for (int idx = tid; idx < n; idx += nthr) tsc[idx] = p.tscale[(size_t)b * n + idx];
  for (int idx = tid; idx < m; idx += nthr) eerr[idx] = p.edge_err[(size_t)b * m + idx];
  // control state of this point
  int iter = p.rs.iter[b], slot = p.rs.slot[b], step = p.rs.step[b], nchecks = p.rs.nchecks[b];
  int conv = 0;
  double last_energy = p.energy[b];
  (void)ctrl;
  __syncthreads();

  // iter / slot / step / nchecks are kept identically by every thread: all decisions below are CTA-uniform
  for (;;) {
    for (int lv = 0; lv < p.n_levels; ++lv) {
      const int e0 = p.level_off[lv], e1 = p.level_off[lv + 1];
      if (e1 - e0 == 1 && p.n_warps >= 2) {
        // a level with ONE edge: its two sides go to warps 0 and 1 (each reduces and recomposes one site tensor, warp 0
        // runs the SVD in between); the other warps only take part in the two phase barriers
        if (warp < 2) {
          fused_edge_update<R, MMAX, CPL, ACCV>(p, descs, wm0, wm, warp + 1, tens, wts, iwts, tsc, eerr, b, p.level_edges[e0], slot, lane);
        } else {
          bar_sync(1, blockDim.x);
          bar_sync(1, blockDim.x);
        }
      } else {
        for (int ei = e0 + warp; ei < e1; ei += p.n_warps)
          fused_edge_update<R, MMAX, CPL, ACCV>(p, descs, wm, wm, 0, tens, wts, iwts, tsc, eerr, b, p.level_edges[ei], slot, lane);
      }
      const long long cb0 = clock64();
      __syncthreads();
      if (p.prof != nullptr && b == 0 && tid == 0) p.prof[5] += clock64() - cb0;
    }
    const long long ce0 = clock64();
    // one sweep done: the reference's check cadence (i % 2 == 0 and i > 0), :132
    const bool check = (iter % 2 == 0 && iter > 0);  // iter is identical in every thread (updated identically below)
    if (check) {
      for (int ek = warp; ek < m; ek += p.n_warps) {
        const double v = fused_pair_expectation(p, descs, wm, tens, wts, tsc, b, ek, lane);
        if (lane == 0) pval[ek] = v;
      }
      __syncthreads();
      if (p.prof != nullptr && b == 0 && tid == 0) p.prof[4] += clock64() - ce0;
    }
    // control flow of run() (:132-188); every thread evaluates it on the same shared values
    step += 1;
    bool advance = false, stop = false;
    if (check) {
      double err = 0.0, en = 0.0;
      bool bad = false;
      for (int e = 0; e < m; ++e) {
        const double v = eerr[e];
        if (v != v) bad = true;
        err = fmax(err, v);
        en += pval[e];
      }
      if (bad) err = nan("");
      en /= n;
      last_energy = en;
      if (tid == 0 && nchecks < p.rs.log_cap) {
        const size_t o = (size_t)b * p.rs.log_cap + nchecks;
        p.rs.log_err[o] = err;
        p.rs.log_energy[o] = en;
        p.rs.log_slot[o] = slot;
        p.rs.log_step[o] = step;
      }
      nchecks += 1;
      if (err <= p.rs.tol && p.rs.is_last[slot]) {
        conv = 1;
        stop = true;
      } else if (err <= p.rs.tol) {
        advance = true;
      }
    }
    if (!stop) {
      if (!advance) {
        iter += 1;
        if (iter >= p.rs.max_iter) advance = true;
      }
      if (advance) {
        iter = 0;
        if (slot + 1 >= p.rs.n_dts)
          stop = true;
        else
          slot += 1;
      }
    }
    __syncthreads();  // everyone has read eerr / pval of this check before the next sweep overwrites them
    if (stop) break;
  }

  // write the point back
  for (int t = 0; t < n; ++t) {
    const SiteDesc &sd = p.sites[t];
    double *dst = (double *)sd.base + (size_t)b * sd.point_stride;
    const int elems = (int)sd.nvirt * p.d;
    const double *src = tens + p.toff[t];
    for (int idx = tid; idx < elems; idx += nthr) dst[idx] = src[idx];
  }
  for (int idx = tid; idx < m * p.dcap; idx += nthr) p.weights[(size_t)b * m * p.dcap + idx] = wts[idx];
  for (int idx = tid; idx < n; idx += nthr) p.tscale[(size_t)b * n + idx] = tsc[idx];
  for (int idx = tid; idx < m; idx += nthr) p.edge_err[(size_t)b * m + idx] = eerr[idx];
  if (p.prof != nullptr && b == 0 && tid == 0) p.prof[7] = step;
  if (tid == 0) {
    p.energy[b] = last_energy;
    p.rs.iter[b] = iter;
    p.rs.slot[b] = slot;
    p.rs.step[b] = step;
    p.rs.nchecks[b] = nchecks;
    p.rs.conv[b] = conv;
    p.rs.active[b] = 0;
  }
}

template <int R, int MMAX, int CPL, bool ACCV>
cudaError_t launch_run_fused_v(const FusedLaunch &p, int smem_bytes, cudaStream_t st, bool dry_run, int *blocks_per_sm) {
  auto kern = run_fused_kernel<R, MMAX, CPL, ACCV>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  if (blocks_per_sm != nullptr) {
    cudaError_t s = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, 32 * p.n_warps, smem_bytes);
    if (s != cudaSuccess) return s;
  }
  if (dry_run) return cudaSuccess;
  kern<<<p.batch, 32 * p.n_warps, smem_bytes, st>>>(p);
  return cudaGetLastError();
}

// svd_rows = rows of theta per lane (template R of the warp SVD); max_m / max_n = the largest bond matrix of the network.
// Narrow bond matrices take a register-resident variant (columns per lane CPL, rows MMAX); everything else the generic one.
// Thetas below 8 x 8 run the accumulating Jacobi without preconditioner (p.svd_accv), all others the V-free iteration.
inline cudaError_t launch_run_fused(const FusedLaunch &p, int svd_rows, int max_m, int max_n, int smem_bytes, cudaStream_t st,
                                    bool dry_run, int *blocks_per_sm) {
  // svd_rows is one of 2, 4, 6, 8, 12, 16 and fixed the shared-memory geometry on the host: the kernel's R must equal it
  const bool dr = dry_run;
  int *bp = blocks_per_sm;
  if (p.svd_accv) {
    if (svd_rows == 2) return launch_run_fused_v<2, 0, 0, true>(p, smem_bytes, st, dr, bp);
    if (svd_rows == 4) return launch_run_fused_v<4, 0, 0, true>(p, smem_bytes, st, dr, bp);
    return cudaErrorInvalidConfiguration;
  }
  switch (svd_rows) {
    case 4:
      if (max_m <= 8 && max_n <= 32) return launch_run_fused_v<4, 8, 1, false>(p, smem_bytes, st, dr, bp);  // star / peps D = 2
      return launch_run_fused_v<4, 0, 0, false>(p, smem_bytes, st, dr, bp);
    case 6:
      if (max_m <= 8 && max_n <= 32) return launch_run_fused_v<6, 8, 1, false>(p, smem_bytes, st, dr, bp);  // star / peps D = 3
      return launch_run_fused_v<6, 0, 0, false>(p, smem_bytes, st, dr, bp);
    case 8:
      if (max_m <= 8 && max_n <= 64) return launch_run_fused_v<8, 8, 2, false>(p, smem_bytes, st, dr, bp);  // peps D = 4 (config 3)
      return launch_run_fused_v<8, 0, 0, false>(p, smem_bytes, st, dr, bp);
    case 12:
      if (max_m <= 16 && max_n <= 128) return launch_run_fused_v<12, 16, 4, false>(p, smem_bytes, st, dr, bp);  // peps D = 5
      return launch_run_fused_v<12, 0, 0, false>(p, smem_bytes, st, dr, bp);
    case 16:
      return launch_run_fused_v<16, 0, 0, false>(p, smem_bytes, st, dr, bp);
    default:
      return cudaErrorInvalidConfiguration;
  }
}
#endif  // RUN_FUSED_IMPL
