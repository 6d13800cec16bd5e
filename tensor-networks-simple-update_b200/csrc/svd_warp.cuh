// svd_warp.cuh — K2 for real float64 thetas of at most 32 x 32: one-sided Jacobi SVD with the whole
// problem resident in REGISTERS of (part of) one warp.
//
// Replaces SimpleUpdate.truncation_svd (reference src/tnsu/simple_update.py:393-409, scipy gesdd) and the
// theta contraction of time_evolution (:476-483) for the shapes of BASELINE configs 1, 2, 3 and 5.
//
// Layout.  A theta with nc (<= 2*NP) columns is handled by NP "processors" of 2 lanes each.  Processor k
// holds two columns x, y of the stacked matrix [A; V] (A = theta or theta^H so that rows >= columns, V the
// accumulated right rotations); lane half h holds rows h*R .. h*R+R-1 of both.  A warp holds 32/(2*NP)
// independent thetas ("groups").  Nothing lives in shared memory during the iteration and there is no
// block barrier: the only communication is
//     - one shfl_xor(1) per dot product (the two row halves), and
//     - an in-place cyclic shift of the x registers by one processor, alternately down and up,
// which realises the odd-even transposition ordering: step A pairs positions (2k, 2k+1), step B pairs
// (2k+1, 2k+2); every rotation writes its outputs swapped, so after n = 2*NP steps every pair of columns
// has met exactly once (the classic odd-even Jacobi ordering, equivalent to the cyclic one).  The idle pair
// of step B (positions n-1 and 0, sitting together in the last processor) gets the "rotation" (c,s)=(0,1),
// which keeps both in place (one changes sign together with its V column: harmless).
//
// Cost per half step and lane at R = 16: 64 SHFL + ~210 FP64 instructions with the accumulator (ACCV); the default
// V-free iteration with carried column norms: 38 SHFL + 117 FP64 of ~290 instructions (measured B200 rates: 1 warp-SHFL
// / clk / SM, 1 warp-DFMA / 2 clk / SM sub-partition; profiles/r01_microbench_b200.txt).  Under load the kernel follows
// the dependent chain of a step (dot -> shuffle -> two rsqrt -> update -> shift) times the 4 warps a scheduler holds, not
// the instruction count: profiles/r02_k2_stall_lines.txt, DESIGN.md section 4.
#pragma once
#include "edge_update.cuh"

#define SVDW_THREADS 128
#define SVDW_MAX_SWEEPS 60
#define SVDW_TRACKED_SWEEPS 12  // V-free iteration: sweeps that carry the column norms along; later ones recompute them per pair

template <int R>
struct SvdwCols {
  double xa[R], xv[R], ya[R], yv[R];
};

// one Jacobi step on the local pair (P, Q) with swapped outputs:  P <- s P + c Q,  Q <- c P - s Q
template <int R>
DEV bool svdw_rotate(double (&pa)[R], double (&pv)[R], double (&qa)[R], double (&qv)[R], bool idle) {
  double al0 = 0.0, al1 = 0.0, be0 = 0.0, be1 = 0.0, ga0 = 0.0, ga1 = 0.0;
#pragma unroll
  for (int r = 0; r < R; r += 2) {
    al0 = fma(pa[r], pa[r], al0);
    be0 = fma(qa[r], qa[r], be0);
    ga0 = fma(pa[r], qa[r], ga0);
    if (r + 1 < R) {
      al1 = fma(pa[r + 1], pa[r + 1], al1);
      be1 = fma(qa[r + 1], qa[r + 1], be1);
      ga1 = fma(pa[r + 1], qa[r + 1], ga1);
    }
  }
  double al = al0 + al1, be = be0 + be1, ga = ga0 + ga1;
  al += __shfl_xor_sync(0xffffffffu, al, 1);
  be += __shfl_xor_sync(0xffffffffu, be, 1);
  ga += __shfl_xor_sync(0xffffffffu, ga, 1);
  const double tol2 = 2.25e-30;  // (1.5e-15)^2, same threshold as the shared-memory kernel
  const double g2 = ga * ga;
  const bool rot = !idle && (g2 > tol2 * al * be) && (g2 > 0.0);
  double c = 1.0, s = 0.0;
  if (rot) {
    const double delta = be - al;
    const double inv_r = rsqrt(fma(delta, delta, 4.0 * g2));
    const double c2 = fma(0.5 * fabs(delta), inv_r, 0.5);
    const double inv_c = rsqrt(c2);
    c = c2 * inv_c;
    s = (delta < 0.0 ? -ga : ga) * inv_r * inv_c;
  }
  if (idle) {
    c = 0.0;
    s = 1.0;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double p0 = pa[r], q0 = qa[r];
    pa[r] = fma(s, p0, c * q0);
    qa[r] = fma(c, p0, -(s * q0));
    const double p1 = pv[r], q1 = qv[r];
    pv[r] = fma(s, p1, c * q1);
    qv[r] = fma(c, p1, -(s * q1));
  }
  return rot;
}

// One step of the odd-even ordering in a form that serves BOTH half steps with the same code (the Jacobi loop body
// then is a single copy of ~300 instructions instead of two): with flip = false it is svdw_rotate on the pair
// (x, y) without the V rows, outputs swapped; with flip = true it is the same on (y, x), written in terms of (x, y):
//   flip = false:  x' = s x + c y,  y' = c x - s y        flip = true:  x' = -s x + c y,  y' = c x + s y
// where (c, s) come from (al, be, ga) = (|P|^2, |Q|^2, P.Q) of the ordered pair (P, Q) = flip ? (y, x) : (x, y).
template <int R>
DEV void svdw_step_a(double (&xa)[R], double (&ya)[R], double &nx, double &ny, int &max_ecos, int &max_esin, bool flip, bool idle, bool exact,
                     bool frozen) {
  // nx, ny: the squared norms of the two columns, CARRIED ALONG (both row halves hold the same values): they are
  // updated from the rotation below instead of being recomputed for every pair — two of the three dot products of a
  // step.  The sweep loop refreshes them from the columns once per sweep; `exact` (warp-uniform: an iteration far
  // beyond its usual length) recomputes them here, as the kernel did before.  Model: tools/experiments/svd_tracked_norms.py
  // (same sweeps, same singular values on oracle thetas, on matrices graded over 30 decades and on rank-deficient ones).
  double ga0 = 0.0, ga1 = 0.0;
#pragma unroll
  for (int r = 0; r < R; r += 2) {
    ga0 = fma(xa[r], ya[r], ga0);
    if (r + 1 < R) ga1 = fma(xa[r + 1], ya[r + 1], ga1);
  }
  double ga = ga0 + ga1;
  ga += __shfl_xor_sync(0xffffffffu, ga, 1);
  if (exact) {
    double nx0 = 0.0, nx1 = 0.0, ny0 = 0.0, ny1 = 0.0;
#pragma unroll
    for (int r = 0; r < R; r += 2) {
      nx0 = fma(xa[r], xa[r], nx0);
      ny0 = fma(ya[r], ya[r], ny0);
      if (r + 1 < R) {
        nx1 = fma(xa[r + 1], xa[r + 1], nx1);
        ny1 = fma(ya[r + 1], ya[r + 1], ny1);
      }
    }
    nx = nx0 + nx1;
    ny = ny0 + ny1;
    nx += __shfl_xor_sync(0xffffffffu, nx, 1);
    ny += __shfl_xor_sync(0xffffffffu, ny, 1);
  }
  const double al = flip ? ny : nx, be = flip ? nx : ny;
  const double tol2 = 2.25e-30;
  const double g2 = ga * ga;
  const double ab = al * be;
  // frozen: this lane's theta has finished while another theta of the warp still iterates — its columns only keep moving
  const bool rot = !idle && !frozen && (g2 > tol2 * ab) && (g2 > 0.0);
  double c = 1.0, s = 0.0;
  if (rot) {
    const double delta = be - al;
    const double inv_r = rsqrt(fma(delta, delta, 4.0 * g2));
    const double c2 = fma(0.5 * fabs(delta), inv_r, 0.5);
    const double inv_c = rsqrt(c2);
    c = c2 * inv_c;
    s = (delta < 0.0 ? -ga : ga) * inv_r * inv_c;
  }
  // binary exponents of cos^2 = g2 / (al be) of the pair and of sin^2 of its rotation (each within 1 of log2), their
  // maxima over the sweep decide whether a confirming sweep is needed (see the sweep loop)
  if (rot) {
    max_ecos = max(max_ecos, ((__double2hiint(g2) >> 20) & 0x7ff) - ((__double2hiint(ab) >> 20) & 0x7ff));
    max_esin = max(max_esin, ((__double2hiint(s * s) >> 20) & 0x7ff) - 1023);
  }
  if (idle) {
    c = 0.0;
    s = 1.0;
  }
  const double sa = flip ? -s : s;  // x' = sa x + c y,  y' = c x - sa y
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double x0 = xa[r], y0 = ya[r];
    xa[r] = fma(sa, x0, c * y0);
    ya[r] = fma(c, x0, -(sa * y0));
  }
  // |x'|^2 = s^2 |x|^2 + c^2 |y|^2 + 2 sa c x.y,   |y'|^2 = c^2 |x|^2 + s^2 |y|^2 - 2 sa c x.y
  {
    const double ss = s * s, cc = c * c, m = 2.0 * c * sa * ga;
    const double nxn = fma(ss, nx, fma(cc, ny, m));
    ny = fma(cc, nx, fma(ss, ny, -m));
    nx = nxn;
  }
}

// Everything one call of the warp SVD needs.  The edge-update kernel K2 (svd_warp_kernel) fills it per (point, edge)
// from global memory; the on-chip run kernel (run_fused.cuh) points it at shared memory.
struct SvdwArgs {
  int NP;           // processors (column pairs) per theta; a theta occupies 2 * NP lanes
  int n_groups;     // thetas per warp handled by this call (lanes beyond n_groups * 2 * NP are surplus)
  bool valid;       // this lane's group has a theta to factor
  int d, Dk, Dnew, Ki, Kj, ni, nj, Mi, Mj;
  double *gs;       // the group's work memory: L_i | L_j | gate | lam_old | theta (| transpose buffer | pivot column)
  int ldm, dcap, d4, precond, ldt, tb_off, max_sweeps;
  const double *src_li, *src_lj;  // L factors [M][ldm] (may equal gs / gs + ldm*ldm: then nothing is copied)
  const double *gate_src;         // d^4 gate of this (point, edge, time step)
  double *lam;       // the edge's weight vector: old values in, lambda' = s / sum(s) out   (reference :296)
  double *edge_err;  // || lambda' - lambda ||_2 of this edge (reference :196-206), NaN when the bond dimension changed
  double *rt_i, *rt_j;  // r~_i, r~_j [d * Dnew][ldm]
  int *status;       // failure counters (EdgeLaunch::status)
  int theta_ready;   // the caller has loaded lam_old and built theta in the group memory already (run_fused.cuh)
  double *dbg_sigma;
  int *dbg_info;     // nullptr unless this theta is the one tnsu_debug_edge asked for
};

// ACCV = true: the right rotations are accumulated in registers next to A (rows [A; V]).
// ACCV = false (needs the QR preconditioner): only A is iterated — 36 % fewer flops, half the shuffles, half the
// registers — and the Dnew right singular vectors that are kept are recovered at the end from the theta that
// stayed in shared memory, w_k = theta^T u_k / ||theta^T u_k||, followed by two passes of modified Gram-Schmidt in
// rank order.  The error of w_k is eps * sigma_1 / sigma_k, and every later use of that vector carries the factor
// lambda_k = sigma_k / sum(sigma), so weights / RDMs / energies keep the 1e-10 parity (checked on all golden cases
// and on product states with weights down to 1e-16 with a numpy model of this route before it was written);
// vectors of (numerically) zero singular values become an arbitrary orthonormal completion, as they are in LAPACK.
// Warp-collective: every lane of the warp must call it (lanes without work pass valid = false).
template <int R, bool ACCV>
DEV void svdw_body(const SvdwArgs &args) {
  const int lane = threadIdx.x & 31;
  const int NP = args.NP;
  const int G = 2 * NP;
  const int gpw = args.n_groups;
  const int grp = lane / G;
  const int gl = lane - grp * G;
  const int k = gl >> 1, h = gl & 1;
  const bool valid = args.valid;
  const int ldm = args.ldm;
  const int lda = 2 * R;
  const bool precond = args.precond != 0;

  double *gs = args.gs;
  double *Li = gs;
  double *Lj = Li + ldm * ldm;
  double *gate_s = Lj + ldm * ldm;
  double *lam_old = gate_s + args.d4;
  double *th = lam_old + args.dcap;

  const int d = args.d, Dk = args.Dk, Dnew = args.Dnew, Ki = args.Ki, Kj = args.Kj;
  // without preconditioner: A = theta or theta^T so that rows >= columns; with it: A' = theta^T always
  const bool transposed = precond ? true : (args.ni < args.nj);
  const int nr = valid ? (transposed ? args.nj : args.ni) : 0;
  const int nc = valid ? (transposed ? args.ni : args.nj) : 0;
  double *lamv = args.lam;
  if (valid && !args.theta_ready) {
    if (args.src_li != Li)
      for (int idx = gl; idx < args.Mi * ldm; idx += G) Li[idx] = args.src_li[idx];
    if (args.src_lj != Lj)
      for (int idx = gl; idx < args.Mj * ldm; idx += G) Lj[idx] = args.src_lj[idx];
    for (int idx = gl; idx < args.d4; idx += G) gate_s[idx] = args.gate_src[idx];
    for (int x = gl; x < Dk; x += G) lam_old[x] = lamv[x];
  }
  if constexpr (!ACCV) {
    // theta stays in shared memory until the end and is read without bounds there: clear the padding
    if (grp < gpw && !args.theta_ready)
      for (int idx = gl; idx < lda * 2 * R; idx += G) th[idx] = 0.0;
  }
  __syncwarp();

  // theta[(qi,i'),(j',qj)] = sum_{i,j,e} L_i[(i,e),qi] lam_e L_j[(j,e),qj] G[i,j,i',j']   (reference :476-483)
  if (valid && !args.theta_ready) {
    for (int idx = gl; idx < Ki * Kj; idx += G) {
      const int qi = idx / Kj, qj = idx - qi * Kj;
      // rolled over (i, j): straight-line code here would only evict the Jacobi loop of the other warps from the
      // instruction cache
      double out[TNSU_MAX_SPIN][TNSU_MAX_SPIN];
#pragma unroll
      for (int ip = 0; ip < TNSU_MAX_SPIN; ++ip)
#pragma unroll
        for (int jp = 0; jp < TNSU_MAX_SPIN; ++jp) out[ip][jp] = 0.0;
#pragma unroll 1
      for (int ij = 0; ij < d * d; ++ij) {
        const int i = ij / d, j = ij - i * d;
        double a = 0.0;
        for (int e = 0; e < Dk; ++e) a = fma(Li[(i * Dk + e) * ldm + qi] * lam_old[e], Lj[(j * Dk + e) * ldm + qj], a);
        const double *gp = gate_s + ij * d * d;
#pragma unroll
        for (int ip = 0; ip < TNSU_MAX_SPIN; ++ip)
#pragma unroll
          for (int jp = 0; jp < TNSU_MAX_SPIN; ++jp)
            if (ip < d && jp < d) out[ip][jp] = fma(a, gp[ip * d + jp], out[ip][jp]);
      }
#pragma unroll
      for (int ip = 0; ip < TNSU_MAX_SPIN; ++ip)
#pragma unroll
        for (int jp = 0; jp < TNSU_MAX_SPIN; ++jp) {
          if (ip < d && jp < d) {
            const double a = out[ip][jp];
            const int row = qi * d + ip, cl = jp * Kj + qj;
            if (transposed)
              th[(size_t)row * lda + cl] = a;  // A = theta^T
            else
              th[(size_t)cl * lda + row] = a;
          }
        }
    }
  }
  __syncwarp();

  const int base = grp * G;
  const bool in_group = grp < gpw;
  const long long clk_a = clock64();

  // my two columns (2k, 2k+1) of [A; V], rows h*R .. h*R+R-1
  constexpr int RV = ACCV ? R : 1;
  double xa[R], xv[RV], ya[R], yv[RV];
  if (ACCV && !precond) {
    const int c0 = 2 * k, c1 = 2 * k + 1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int gr = h * R + r;
      xa[r] = (valid && gr < nr && c0 < nc) ? th[(size_t)c0 * lda + gr] : 0.0;
      ya[r] = (valid && gr < nr && c1 < nc) ? th[(size_t)c1 * lda + gr] : 0.0;
      if constexpr (ACCV) {
        xv[r] = (valid && gr == c0 && c0 < nc) ? 1.0 : 0.0;
        yv[r] = (valid && gr == c1 && c1 < nc) ? 1.0 : 0.0;
      }
    }
  } else {
    // ---- preconditioner (Drmac-Veselic): Householder QR with column pivoting A' P = Q R, then Jacobi runs on
    // X = R~^T (R~ = R P^T: columns never move, so no permutation has to be undone) with the accumulator
    // started at Q:  X W = Y  =>  A' = (Q W) (Y)^T, i.e. left vectors of A' = final accumulator, right vectors =
    // normalised final X columns.  Measured on the oracle's thetas: 13.3 -> 7.0 Jacobi sweeps.
    // One column of A' and of Z = Q^T per lane (rows in registers); the pivot column travels through shared
    // memory (broadcast reads), the pivot search is one redux over packed (norm, lane) keys.
    const int ldt = args.ldt;
    // ACCV: th is reused as the 2R x ldt transpose buffer once A' sits in registers.  !ACCV: theta must survive, the
    // (half-size, R x ldt) buffer and the pivot column live at svdw_tb_off (on the dead L_i / L_j when they fit)
    double *tb = ACCV ? th : gs + args.tb_off;
    double *wbuf = tb + (size_t)(ACCV ? 2 * R : R) * ldt;  // [2R + 1] pivot column and its trailing norm
    constexpr int RZ = ACCV ? 2 * R : 1;
    double A[2 * R], Z[RZ];
    const int cj = gl;
    const bool cvalid = valid && in_group && cj < nc;
    double cn = 0.0;
#pragma unroll
    for (int r = 0; r < 2 * R; ++r) {
      A[r] = (cvalid && r < nr) ? th[(size_t)cj * lda + r] : 0.0;
      if constexpr (ACCV) Z[r] = (valid && in_group && r == cj && cj < nr) ? 1.0 : 0.0;
      cn = fma(A[r], A[r], cn);
    }
    __syncwarp();
    const int nsteps = valid ? min(nr, nc) : 0;
    // surplus lanes (beyond the last group) share ONE mask: a mask per lane would make the warp execute the pivot
    // search once per distinct mask (measured in the on-chip run kernel: 16 surplus lanes = 17 serialised redux per step)
    const unsigned all_groups = (gpw * G >= 32) ? 0xffffffffu : ((1u << (gpw * G)) - 1u);
    const unsigned gmask = in_group ? (G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << base)) : ~all_groups;
    bool done = !cvalid;
    const int max_steps = min(2 * R, G);
#pragma unroll 1
    for (int j = 0; j < max_steps; ++j) {
      const unsigned key = (!done && j < nsteps) ? (((unsigned)__double2hiint(cn) & ~63u) | (unsigned)(63 - cj)) : 0u;
      const unsigned best = __reduce_max_sync(gmask, key);
      const bool have = best != 0u;
      if (!__any_sync(0xffffffffu, have)) break;
      const int pl = 63 - (int)(best & 63u);
      if (have && cj == pl) {
        done = true;
#pragma unroll
        for (int r = 0; r < 2 * R; ++r) wbuf[r] = A[r];
        wbuf[2 * R] = cn;  // the norm of its trailing part (rows >= j) was accumulated by the previous reflection
      }
      __syncwarp();
      const double alpha = wbuf[j];
      const double n2 = wbuf[2 * R];
      double tau = 0.0, head = alpha;
      if (have && n2 > 0.0) {
        const double inv = rsqrt(n2);
        const double norm = n2 * inv;
        const double sgn = (alpha < 0.0) ? -1.0 : 1.0;
        head = fma(sgn, norm, alpha);
        tau = 1.0 / fma(norm, fabs(alpha), n2);
      }
      __syncwarp();
      if (gl == 0 && in_group) wbuf[j] = head;
      __syncwarp();
      // pass 1: w^T a for my columns of A' and Z.  4-row blocks entirely above row j are skipped (j is warp
      // uniform), only the block that contains row j needs its upper rows masked
      double dA, dZ;
      {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
#pragma unroll
        for (int r0 = 0; r0 < 2 * R; r0 += 4) {
          if (r0 + 3 >= j) {
            double w0 = wbuf[r0], w1 = wbuf[r0 + 1], w2 = wbuf[r0 + 2];
            const double w3 = wbuf[r0 + 3];
            if (r0 < j) {
              w0 = 0.0;
              w1 = (r0 + 1 >= j) ? w1 : 0.0;
              w2 = (r0 + 2 >= j) ? w2 : 0.0;
            }
            a0 = fma(w0, A[r0], a0);
            a1 = fma(w1, A[r0 + 1], a1);
            a2 = fma(w2, A[r0 + 2], a2);
            a3 = fma(w3, A[r0 + 3], a3);
            if constexpr (ACCV) {
              z0 = fma(w0, Z[r0], z0);
              z1 = fma(w1, Z[r0 + 1], z1);
              z2 = fma(w2, Z[r0 + 2], z2);
              z3 = fma(w3, Z[r0 + 3], z3);
            }
          }
        }
        dA = (a0 + a1) + (a2 + a3);
        dZ = (z0 + z1) + (z2 + z3);
      }
      const double fA = tau * dA, fZ = tau * dZ;
      // pass 2: reflect; trailing column norms (rows > j) for the next pivot search and the next reflector
      double cnn;
      {
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
        for (int r0 = 0; r0 < 2 * R; r0 += 4) {
          if (r0 + 3 >= j) {
            double w0 = wbuf[r0], w1 = wbuf[r0 + 1], w2 = wbuf[r0 + 2];
            const double w3 = wbuf[r0 + 3];
            if (r0 <= j) {  // the block that contains row j: rows above it untouched, row j itself not counted
              w0 = (r0 >= j) ? w0 : 0.0;
              w1 = (r0 + 1 >= j) ? w1 : 0.0;
              w2 = (r0 + 2 >= j) ? w2 : 0.0;
              A[r0] = fma(-fA, w0, A[r0]);
              A[r0 + 1] = fma(-fA, w1, A[r0 + 1]);
              A[r0 + 2] = fma(-fA, w2, A[r0 + 2]);
              A[r0 + 3] = fma(-fA, w3, A[r0 + 3]);
              c1 = (r0 + 1 > j) ? fma(A[r0 + 1], A[r0 + 1], c1) : c1;
              c2 = (r0 + 2 > j) ? fma(A[r0 + 2], A[r0 + 2], c2) : c2;
              c3 = (r0 + 3 > j) ? fma(A[r0 + 3], A[r0 + 3], c3) : c3;
            } else {
              A[r0] = fma(-fA, w0, A[r0]);
              A[r0 + 1] = fma(-fA, w1, A[r0 + 1]);
              A[r0 + 2] = fma(-fA, w2, A[r0 + 2]);
              A[r0 + 3] = fma(-fA, w3, A[r0 + 3]);
              c0 = fma(A[r0], A[r0], c0);
              c1 = fma(A[r0 + 1], A[r0 + 1], c1);
              c2 = fma(A[r0 + 2], A[r0 + 2], c2);
              c3 = fma(A[r0 + 3], A[r0 + 3], c3);
            }
            if constexpr (ACCV) {
              Z[r0] = fma(-fZ, w0, Z[r0]);
              Z[r0 + 1] = fma(-fZ, w1, Z[r0 + 1]);
              Z[r0 + 2] = fma(-fZ, w2, Z[r0 + 2]);
              Z[r0 + 3] = fma(-fZ, w3, Z[r0 + 3]);
            }
          }
        }
        cnn = (c0 + c1) + (c2 + c3);
      }
      cn = cnn;
      __syncwarp();
    }
    // X = R~^T and accumulator = Q (thin), through the transpose buffer: writer (row r, column cj) -> tb[r][cj],
    // reader processor k slot s, row c = h*R + r' wants R~[2k+s][c] / Z[2k+s][c]
    const int j0 = 2 * k, j1 = 2 * k + 1;
    if constexpr (ACCV) {
      if (in_group) {
#pragma unroll
        for (int r = 0; r < 2 * R; ++r) tb[(size_t)r * ldt + cj] = A[r];
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int c = h * R + r;
        xa[r] = (valid && c < nc && j0 < nsteps) ? tb[(size_t)j0 * ldt + c] : 0.0;
        ya[r] = (valid && c < nc && j1 < nsteps) ? tb[(size_t)j1 * ldt + c] : 0.0;
      }
      __syncwarp();
      if (in_group) {
#pragma unroll
        for (int r = 0; r < 2 * R; ++r) tb[(size_t)r * ldt + cj] = Z[r];
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int c = h * R + r;
        xv[r] = (valid && c < nr && j0 < nsteps) ? tb[(size_t)j0 * ldt + c] : 0.0;
        yv[r] = (valid && c < nr && j1 < nsteps) ? tb[(size_t)j1 * ldt + c] : 0.0;
      }
    } else {
      // rows 0..R-1 of R~ first (read by the processors whose columns 2k, 2k+1 lie there), then rows R..2R-1
#pragma unroll
      for (int r = 0; r < R; ++r) xa[r] = ya[r] = 0.0;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        __syncwarp();
        if (in_group) {
#pragma unroll
          for (int r = 0; r < R; ++r) tb[(size_t)r * ldt + cj] = A[half * R + r];
        }
        __syncwarp();
        if (j0 / R == half) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int c = h * R + r;
            xa[r] = (valid && c < nc && j0 < nsteps) ? tb[(size_t)(j0 - half * R) * ldt + c] : 0.0;
            ya[r] = (valid && c < nc && j1 < nsteps) ? tb[(size_t)(j1 - half * R) * ldt + c] : 0.0;
          }
        }
      }
    }
  }

  // source lanes of the cyclic x shifts (invalid / surplus lanes shuffle with themselves)
  const int src_next = in_group ? base + 2 * ((k + 1 == NP) ? 0 : k + 1) + h : lane;  // layout A -> B
  const int src_prev = in_group ? base + 2 * ((k == 0) ? NP - 1 : k - 1) + h : lane;  // layout B -> A
  const bool last = (k == NP - 1);

  const long long clk_b = clock64();
  int sweeps = 0;
  bool jacobi_done = false;
  const int max_sweeps = min(args.max_sweeps, SVDW_MAX_SWEEPS);
  if constexpr (ACCV) {
    for (; sweeps < max_sweeps; ++sweeps) {
      bool rotated = false;
      for (int st = 0; st < NP; ++st) {
        // step A: pairs (x_k, y_k)
        rotated |= svdw_rotate<R>(xa, xv, ya, yv, false);
        // layout A -> B: x_k <- x_{k+1}  (the last processor receives position 0)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          xa[r] = __shfl_sync(0xffffffffu, xa[r], src_next);
          xv[r] = __shfl_sync(0xffffffffu, xv[r], src_next);
        }
        // step B: pairs (y_k, x_k) = positions (2k+1, 2k+2); the last processor idles
        rotated |= svdw_rotate<R>(ya, yv, xa, xv, last);
        // layout B -> A
#pragma unroll
        for (int r = 0; r < R; ++r) {
          xa[r] = __shfl_sync(0xffffffffu, xa[r], src_prev);
          xv[r] = __shfl_sync(0xffffffffu, xv[r], src_prev);
        }
      }
      if (!__any_sync(0xffffffffu, rotated)) {
        ++sweeps;
        jacobi_done = true;
        break;
      }
    }
  } else {
    // one half step per trip (step A: pair (x_k, y_k), then x_k <- x_{k+1}; step B: pair (y_k, x_k), the last
    // processor idle, then x_k <- x_{k-1}): a single copy of the step in the instruction stream
    // The iteration ends after a sweep without rotation — or after a sweep whose rotations were all SMALL, in the sense
    // (max cosine of a pair met) x (max |sin| of a rotation applied) <= 1.6e-15: a pair made orthogonal in such a sweep is
    // disturbed afterwards only by the later rotations of its two columns, each adding at most |sin| x (cosine of the
    // third column), at most 2 (n - 1) = 62 times at n = 32: the cosines left behind are <= 1e-13, the singular values
    // (second order in them) exact to rounding and the vectors good to 1e-13 — the confirming sweep, one of the ~7 of a
    // 32 x 32 theta, would only establish that.  Measured on oracle thetas (tools/experiments/svd_tracked_norms.py): 0.75
    // to 0.9 sweeps saved per theta, cosines left behind <= 1.6e-15.  Pairs of (nearly) equal singular values rotate by
    // large angles at tiny cosines: they do not pass, and the usual rotation-free sweep confirms them.
    // The test runs on binary exponents: log2(cos^2 sin^2) < max_ecos + max_esin + 2 <= -99 < log2(2.6e-30).
    // The decision is taken PER THETA (a warp hosts up to four small ones): a finished theta is frozen — its columns keep
    // moving with the others' but are not rotated any more — so what a point computes does not depend on which other
    // points share its warp (tests: test_launches_over_the_running_points_only compares bit for bit).
    const unsigned jall_groups = (gpw * G >= 32) ? 0xffffffffu : ((1u << (gpw * G)) - 1u);
    const unsigned jmask = in_group ? (G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << base)) : ~jall_groups;
    bool frozen = !valid;
    int src_pending = lane;
    for (; sweeps < max_sweeps; ++sweeps) {
      int max_ecos = -100000, max_esin = -100000;  // no rotation yet
      // the column norms travel with the columns and are refreshed from them once per sweep (svdw_step_a)
      double nx, ny;
      {
        double nx0 = 0.0, nx1 = 0.0, ny0 = 0.0, ny1 = 0.0;
#pragma unroll
        for (int r = 0; r < R; r += 2) {
          nx0 = fma(xa[r], xa[r], nx0);
          ny0 = fma(ya[r], ya[r], ny0);
          if (r + 1 < R) {
            nx1 = fma(xa[r + 1], xa[r + 1], nx1);
            ny1 = fma(ya[r + 1], ya[r + 1], ny1);
          }
        }
        nx = nx0 + nx1;
        ny = ny0 + ny1;
        nx += __shfl_xor_sync(0xffffffffu, nx, 1);
        ny += __shfl_xor_sync(0xffffffffu, ny, 1);
      }
      const bool exact = sweeps >= SVDW_TRACKED_SWEEPS;
#pragma unroll 1
      for (int st = 0; st < 2 * NP; ++st) {
        // the cyclic shift that follows a step is applied at the top of the NEXT one: the shuffled values are consumed
        // by the rotation right away and its results land in the loop-carried registers (no register copies at the
        // loop end).  The shift still pending when the iteration ends only decides which lane holds which column.
#pragma unroll
        for (int r = 0; r < R; ++r) xa[r] = __shfl_sync(0xffffffffu, xa[r], src_pending);
        nx = __shfl_sync(0xffffffffu, nx, src_pending);
        const bool flip = (st & 1) != 0;
        svdw_step_a<R>(xa, ya, nx, ny, max_ecos, max_esin, flip, flip && last, exact, frozen);
        src_pending = flip ? src_prev : src_next;
      }
      if (__reduce_max_sync(jmask, max_ecos) + __reduce_max_sync(jmask, max_esin) <= -101) frozen = true;
      if (__all_sync(0xffffffffu, frozen)) {
        ++sweeps;
        jacobi_done = true;
        break;
      }
    }
  }

  const long long clk_c = clock64();
  // singular values = column norms; dummy (padding) columns have V == 0 and sort last
  double sx, sy;
  {
    double nx = 0.0, ny = 0.0, vx = 0.0, vy = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      nx = fma(xa[r], xa[r], nx);
      ny = fma(ya[r], ya[r], ny);
      if constexpr (ACCV) {
        vx = fma(xv[r], xv[r], vx);
        vy = fma(yv[r], yv[r], vy);
      }
    }
    nx += __shfl_xor_sync(0xffffffffu, nx, 1);
    ny += __shfl_xor_sync(0xffffffffu, ny, 1);
    if constexpr (ACCV) {
      vx += __shfl_xor_sync(0xffffffffu, vx, 1);
      vy += __shfl_xor_sync(0xffffffffu, vy, 1);
      sx = (vx > 0.5) ? sqrt(nx) : -1.0;
      sy = (vy > 0.5) ? sqrt(ny) : -1.0;
    } else {
      // padding columns are exactly zero and stay so: they tie with genuine zero columns, which are equivalent
      sx = sqrt(nx);
      sy = sqrt(ny);
    }
  }
  // rank of my columns in descending order (ties: lower (processor, slot) first)
  int rank_x = 0, rank_y = 0;
  for (int j = 0; j < NP; ++j) {
    const int src = in_group ? base + 2 * j + h : lane;
    const double ox = __shfl_sync(0xffffffffu, sx, src);
    const double oy = __shfl_sync(0xffffffffu, sy, src);
    rank_x += (ox > sx || (ox == sx && 2 * j < 2 * k)) ? 1 : 0;
    rank_x += (oy > sx || (oy == sx && 2 * j + 1 < 2 * k)) ? 1 : 0;
    rank_y += (ox > sy || (ox == sy && 2 * j < 2 * k + 1)) ? 1 : 0;
    rank_y += (oy > sy || (oy == sy && 2 * j + 1 < 2 * k + 1)) ? 1 : 0;
  }
  const bool sel_x = valid && rank_x < Dnew;
  const bool sel_y = valid && rank_y < Dnew;
  // sum(s) and the convergence error are summed IN RANK ORDER (through a dead buffer of at least 2 R doubles): which
  // lane holds which column at the end depends on how many sweeps the warp ran — with several small thetas per warp on
  // the slowest of them — and sums in lane order would make a point's last bits depend on its neighbours in the batch
  double *rs = ACCV ? th : gs + args.tb_off;  // ACCV: theta is dead; V-free: the transpose buffer, free until the recovery
  const int dn = valid ? Dnew : 0;
  if (sel_x && h == 0) rs[rank_x] = sx;
  if (sel_y && h == 0) rs[rank_y] = sy;
  __syncwarp();
  double tot = 0.0;
  for (int r = 0; r < dn; ++r) tot += rs[r];
  __syncwarp();
  if (valid && gl == 0 && !(tot > 0.0 && tot < 1e300)) atomicOr(args.status, 1);  // NaN / Inf in the state (or theta == 0)
  if (!jacobi_done && lane == 0) atomicAdd(args.status + 1, 1);
  // lambda' = s / sum(s)  (reference :296) and this edge's convergence error (:196-206)
  if (sel_x) {
    const double ln = sx / tot;
    const double df = ln - lam_old[rank_x];
    if (h == 0) {
      lamv[rank_x] = ln;
      rs[rank_x] = df * df;
    }
  }
  if (sel_y) {
    const double ln = sy / tot;
    const double df = ln - lam_old[rank_y];
    if (h == 0) {
      lamv[rank_y] = ln;
      rs[rank_y] = df * df;
    }
  }
  __syncwarp();
  double err2 = 0.0;
  for (int r = 0; r < dn; ++r) err2 += rs[r];
  if (valid && gl == 0) *args.edge_err = (Dnew == Dk) ? sqrt(err2) : nan("");

  if constexpr (!ACCV) {
    // ---- right singular vectors of the Dnew columns that are kept, from the theta left in shared memory ----
    // u = a / sigma in place (left vectors, theta-row index c = h*R + r)
    {
      const double ix = (sel_x && sx > 0.0) ? 1.0 / sx : 0.0, iy = (sel_y && sy > 0.0) ? 1.0 / sy : 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        xa[r] *= ix;
        ya[r] *= iy;
      }
    }
    // w_j[c'] = sum_c theta[c][c'] u_j[c] = sum_c th[c*lda + c'] u_j[c] for the KEPT columns j only (Dnew of the 2 NP the
    // processors hold): the u_j go to the dead transpose buffer in rank order, lane gl then owns the output index
    // c' = gl (, gl + G, ...) of EVERY kept column — one theta element loaded per 8 multiply-adds, 8 independent
    // accumulators — and the results return to the processor layout (wx / wy of the lanes that hold column j) through
    // the same buffer.  (Before: every processor computed both of its columns, kept or not: 4x the loads and FMAs at
    // Dnew = 8 of 32, 9 % of the kernel's stall samples.)  Kept columns are taken SVDW_JB at a time, as many as the
    // buffer holds next to their results.  A lane's u_j is dead once it sits in the buffer: w_j takes its registers (the
    // columns that are not kept were zeroed above and stay zero), so A and W never sit in registers together.
    double *wt = gs + args.tb_off;  // R * ldt + 2 R + 2 doubles, free since the transposition
    double *rt_i = valid ? args.rt_i : nullptr;
    const int n_rowidx = valid ? args.ni : 0;
    // theta-row index (qi, i') of my first row, advanced without divisions
    const int gr0 = h * R;
    const int qi0 = gr0 / d, ip0 = gr0 - qi0 * d;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      if (which ? sel_y : sel_x) {
        const int xp = which ? rank_y : rank_x;
        int qi = qi0, ip = ip0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (gr0 + r < n_rowidx) rt_i[(ip * Dnew + xp) * ldm + qi] = which ? ya[r] : xa[r];
          if (++ip == d) {
            ip = 0;
            ++qi;
          }
        }
      }
    }
    {
      constexpr int JBMAX = 8;
      const int jb = max(1, min(JBMAX, (R * args.ldt + 2 * R + 2) / (4 * R)));
      double *usm = wt;                 // [2R][jb]  u_j[c] of the current batch
      double *wsm = wt + 2 * R * jb;    // [jb][2R]  w_j[c']
      const int jall = __reduce_max_sync(0xffffffffu, valid ? Dnew : 0);
      const bool writer = valid && in_group;
#pragma unroll 1
      for (int j0 = 0; j0 < jall; j0 += jb) {
        const int jn = valid ? max(0, min(jb, Dnew - j0)) : 0;
        const bool bx = sel_x && rank_x >= j0 && rank_x < j0 + jb, by = sel_y && rank_y >= j0 && rank_y < j0 + jb;
        if (bx) {
#pragma unroll
          for (int r = 0; r < R; ++r) usm[(h * R + r) * jb + (rank_x - j0)] = xa[r];
        }
        if (by) {
#pragma unroll
          for (int r = 0; r < R; ++r) usm[(h * R + r) * jb + (rank_y - j0)] = ya[r];
        }
        __syncwarp();
#pragma unroll 1
        for (int cp = gl; cp < 2 * R; cp += G) {
          double acc[JBMAX];
#pragma unroll
          for (int j = 0; j < JBMAX; ++j) acc[j] = 0.0;
          const double *tcol = th + cp;
#pragma unroll 4
          for (int c = 0; c < 2 * R; ++c) {
            const double t = tcol[(size_t)c * lda];
            const double *uc = usm + c * jb;
#pragma unroll
            for (int j = 0; j < JBMAX; ++j)
              if (j < jn) acc[j] = fma(t, uc[j], acc[j]);
          }
          if (writer) {
#pragma unroll
            for (int j = 0; j < JBMAX; ++j)
              if (j < jn) wsm[j * 2 * R + cp] = acc[j];
          }
        }
        __syncwarp();
        if (bx) {
#pragma unroll
          for (int r = 0; r < R; ++r) xa[r] = wsm[(rank_x - j0) * 2 * R + h * R + r];
        }
        if (by) {
#pragma unroll
          for (int r = 0; r < R; ++r) ya[r] = wsm[(rank_y - j0) * 2 * R + h * R + r];
        }
        __syncwarp();
      }
    }
    double(&wx)[R] = xa, (&wy)[R] = ya;  // from here on the registers hold the right vectors of the kept columns
    // a kept column with no direction of its own (sigma = 0, or theta^T u underflows) starts from a fixed dense vector
    {
      double n2x = 0.0, n2y = 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        n2x = fma(wx[r], wx[r], n2x);
        n2y = fma(wy[r], wy[r], n2y);
      }
      n2x += __shfl_xor_sync(0xffffffffu, n2x, 1);
      n2y += __shfl_xor_sync(0xffffffffu, n2y, 1);
      const bool fx = sel_x && !(n2x > 1e-280), fy = sel_y && !(n2y > 1e-280);
      if (__any_sync(0xffffffffu, fx || fy))  // rare: keep it out of the instruction stream otherwise
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int c = h * R + r;
        if (fx) wx[r] = (c < nr) ? (double)(((unsigned)(c + 1) * 2654435761u ^ (unsigned)(rank_x + 1) * 40503u) >> 8 & 0xffffu) - 32767.5 : 0.0;
        if (fy) wy[r] = (c < nr) ? (double)(((unsigned)(c + 1) * 2654435761u ^ (unsigned)(rank_y + 1) * 40503u) >> 8 & 0xffffu) - 32767.5 : 0.0;
      }
    }
    // two passes of modified Gram-Schmidt in rank order; the owner of rank j normalises, broadcasts, the rest project
    const unsigned gbits = in_group ? (G >= 32 ? 0xffffffffu : (((1u << G) - 1u) << base)) : 0u;
    const int jmax = __reduce_max_sync(0xffffffffu, valid ? Dnew : 0);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
      for (int j = 0; j < jmax; ++j) {
        const bool own_x = sel_x && rank_x == j, own_y = sel_y && rank_y == j;
        double n2x = 0.0, n2y = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          n2x = fma(wx[r], wx[r], n2x);
          n2y = fma(wy[r], wy[r], n2y);
        }
        n2x += __shfl_xor_sync(0xffffffffu, n2x, 1);
        n2y += __shfl_xor_sync(0xffffffffu, n2y, 1);
        const double scx = (own_x && n2x > 0.0) ? rsqrt(n2x) : 1.0, scy = (own_y && n2y > 0.0) ? rsqrt(n2y) : 1.0;
        const unsigned bo = (__ballot_sync(0xffffffffu, own_x || own_y)) & gbits;
        const int src = bo ? ((__ffs(bo) - 1) | h) : lane;
        double q[R];
        double dx = 0.0, dy = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          wx[r] *= scx;
          wy[r] *= scy;
          q[r] = __shfl_sync(0xffffffffu, own_y ? wy[r] : wx[r], src);
          dx = fma(wx[r], q[r], dx);
          dy = fma(wy[r], q[r], dy);
        }
        dx += __shfl_xor_sync(0xffffffffu, dx, 1);
        dy += __shfl_xor_sync(0xffffffffu, dy, 1);
        const double px = (bo && sel_x && rank_x > j) ? dx : 0.0, py = (bo && sel_y && rank_y > j) ? dy : 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          wx[r] = fma(-px, q[r], wx[r]);
          wy[r] = fma(-py, q[r], wy[r]);
        }
      }
    }
    if (valid) {
      double *rt_j = args.rt_j;
      const int n_colidx = args.nj;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const bool sel = which ? sel_y : sel_x;
        const int xp = which ? rank_y : rank_x;
        if (sel) {
          int jp = (h * R) / Kj, qj = h * R - jp * Kj;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (h * R + r < n_colidx) rt_j[(jp * Dnew + xp) * ldm + qj] = which ? wy[r] : wx[r];
            if (++qj == Kj) {
              qj = 0;
              ++jp;
            }
          }
        }
      }
    }
  }

  // r~_i[(i',x'),qi] = U[(qi,i'),x'],   r~_j[(j',x'),qj] = V^T[x',(j',qj)]     (reference :268-271)
  if (valid) {
    double *rt_i = args.rt_i, *rt_j = args.rt_j;
    // which register half carries the theta-ROW-index vector: A/sigma unless A = theta^T was iterated directly;
    // with the preconditioner A' = theta^T but the roles are swapped once more (X = R^T), so it is A/sigma again
    const bool row_is_v = transposed && !precond;
    const int n_rowidx = args.ni, n_colidx = args.nj;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const bool sel = ACCV && (which ? sel_y : sel_x);
      const int xp = which ? rank_y : rank_x;
      const double sv = which ? sy : sx;
      if (sel) {
        const double isv = sv > 0.0 ? 1.0 / sv : 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int gr = h * R + r;
          const double av = (which ? ya[r] : xa[r]) * isv;  // U (or V of theta when transposed) row gr
          const double vv = ACCV ? (which ? yv[r] : xv[r]) : 0.0;
          // non-transposed: A rows are theta rows (qi,i'), V rows are theta columns (j',qj); transposed: the reverse
          const double row_val = row_is_v ? vv : av;  // value attached to theta ROW index gr
          const double col_val = row_is_v ? av : vv;  // value attached to theta COLUMN index gr
          if (gr < n_rowidx) {
            const int qi = gr / d, ip = gr - qi * d;
            rt_i[(ip * Dnew + xp) * ldm + qi] = row_val;
          }
          if (gr < n_colidx) {
            const int jp = gr / Kj, qj = gr - jp * Kj;
            rt_j[(jp * Dnew + xp) * ldm + qj] = col_val;
          }
        }
      }
    }
    if (args.dbg_info != nullptr) {
      if (sel_x && h == 0) args.dbg_sigma[rank_x] = sx;
      if (sel_y && h == 0) args.dbg_sigma[rank_y] = sy;
      if (!sel_x && h == 0 && rank_x < TNSU_MAX_SVD && sx >= 0.0) args.dbg_sigma[rank_x] = sx;
      if (!sel_y && h == 0 && rank_y < TNSU_MAX_SVD && sy >= 0.0) args.dbg_sigma[rank_y] = sy;
      if (gl == 0) {
        args.dbg_info[0] = args.Mi;
        args.dbg_info[1] = Ki;
        args.dbg_info[2] = args.Mj;
        args.dbg_info[3] = Kj;
        args.dbg_info[4] = args.ni;
        args.dbg_info[5] = args.nj;
        args.dbg_info[6] = Dnew;
        args.dbg_info[7] = sweeps;
        args.dbg_info[8] = (int)(clk_b - clk_a);   // load / QR preconditioner
        args.dbg_info[9] = (int)(clk_c - clk_b);   // Jacobi
        args.dbg_info[10] = (int)(clock64() - clk_c);  // extract
      }
    }
  }
}

#ifndef RUN_FUSED_IMPL  // the on-chip run kernel's translation unit only needs svdw_body
template <int R, bool ACCV>
__global__ void __launch_bounds__(SVDW_THREADS, (128 / SVDW_THREADS) * (ACCV ? ((R > 12) ? 3 : ((R > 6) ? 4 : 6)) : ((R > 8) ? 4 : 6)))
    svd_warp_kernel(const EdgeLaunch<double> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int G = 2 * p.svdw_np;
  const int gpw = 32 / G;  // groups per warp
  const int grp = lane / G;
  const int n_items = p.batch * p.n_level_edges;
  const int item = (blockIdx.x * (SVDW_THREADS / 32) + warp) * gpw + grp;
  bool valid = (grp < gpw) && (item < n_items);
  int b = 0, ek = 0;
  if (valid) {
    const int bi = item / p.n_level_edges;
    ek = p.level_edges[item - bi * p.n_level_edges];
    b = p.point_list != nullptr ? p.point_list[bi] : bi;
    if (p.active != nullptr && p.active[b] == 0) valid = false;
  }
  // whole-warp early out (no group of this warp has work)
  if (!__any_sync(0xffffffffu, valid)) return;
  SvdwArgs a{};
  a.NP = p.svdw_np;
  a.n_groups = gpw;
  a.valid = valid;
  a.gs = (double *)smem_raw + (size_t)(warp * gpw + (grp < gpw ? grp : 0)) * p.svdw_group_doubles;
  a.ldm = p.ldm;
  a.dcap = p.dcap;
  a.d4 = p.d4;
  a.precond = p.svdw_precond;
  a.ldt = p.svdw_ldt;
  a.tb_off = p.svdw_tb_off;
  a.max_sweeps = p.svd_max_sweeps;
  a.status = p.status;
  if (valid) {
    const EdgeDesc &ed = p.descs[ek];
    a.d = ed.d;
    a.Dk = ed.Dk;
    a.Dnew = ed.Dnew;
    a.Ki = ed.s[0].K;
    a.Kj = ed.s[1].K;
    a.ni = ed.ni;
    a.nj = ed.nj;
    a.Mi = ed.s[0].M;
    a.Mj = ed.s[1].M;
    const int ldm2 = p.ldm * p.ldm;
    double *small = p.small + (size_t)item * SM_COUNT * ldm2;
    a.src_li = small + SM_L0 * ldm2;
    a.src_lj = small + SM_L1 * ldm2;
    a.rt_i = small + SM_R0 * ldm2;
    a.rt_j = small + SM_R1 * ldm2;
    const int slot = p.point_slot ? p.point_slot[b] : p.fixed_slot;
    a.gate_src = p.gates + (size_t)slot * p.gate_slot_stride + ((size_t)b * p.m_edges + ek) * p.d4;
    a.lam = p.weights + ((size_t)b * p.m_edges + ek) * p.dcap;
    a.edge_err = p.edge_err + (size_t)b * p.m_edges + ek;
    if (p.dbg_info != nullptr && b == p.dbg_point) {
      a.dbg_sigma = p.dbg_sigma;
      a.dbg_info = p.dbg_info;
    }
  }
  svdw_body<R, ACCV>(a);
}

template <int R, bool ACCV>
cudaError_t launch_svd_warp_r(const EdgeLaunch<double> &p, int n_items, cudaStream_t st) {
  auto kern = svd_warp_kernel<R, ACCV>;
  if (cudaError_t s = tnsu_opt_in_smem((const void *)kern); s != cudaSuccess) return s;
  const int gpw = 32 / (2 * p.svdw_np);
  const int per_cta = (SVDW_THREADS / 32) * gpw;
  const int smem = per_cta * p.svdw_group_doubles * 8;
  kern<<<(n_items + per_cta - 1) / per_cta, SVDW_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

// rows per lane -> kernel variant; p.svdw_rows = ceil(nr_max / 2)
inline cudaError_t launch_svd_warp(const EdgeLaunch<double> &p, int n_items, cudaStream_t st) {
  const int r = p.svdw_rows;
  if (p.svdw_precond && !p.svdw_accv) {  // the V-free iteration exists with the preconditioner only (rows >= 8)
    if (r <= 4) return launch_svd_warp_r<4, false>(p, n_items, st);
    if (r <= 6) return launch_svd_warp_r<6, false>(p, n_items, st);
    if (r <= 8) return launch_svd_warp_r<8, false>(p, n_items, st);
    if (r <= 12) return launch_svd_warp_r<12, false>(p, n_items, st);
    return launch_svd_warp_r<16, false>(p, n_items, st);
  }
  if (r <= 2) return launch_svd_warp_r<2, true>(p, n_items, st);
  if (r <= 4) return launch_svd_warp_r<4, true>(p, n_items, st);
  if (r <= 6) return launch_svd_warp_r<6, true>(p, n_items, st);
  if (r <= 8) return launch_svd_warp_r<8, true>(p, n_items, st);
  if (r <= 12) return launch_svd_warp_r<12, true>(p, n_items, st);
  return launch_svd_warp_r<16, true>(p, n_items, st);
}
#endif  // !RUN_FUSED_IMPL
