// common.cuh — scalar algebra (float64 / complex128), warp helpers and the edge descriptors shared by
// the Simple-Update kernels.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TNSU_MAX_LEGS 8   // max coordination number z of a site tensor
#define TNSU_MAX_SPIN 4   // max physical dimension d
#define TNSU_MAX_M 32     // max d*D rows of a bond matrix (one lane per row in the reductions)
#define TNSU_MAX_SVD 64   // max rows/cols of theta
#define TNSU_MAX_WARPS 16 // max warps per side group

#define TNSU_SMEM_OPT_IN (227 * 1024)

// Opt a kernel in to the full dynamic shared memory of the device the caller is on.  The attribute is per DEVICE, an
// engine may sit on any of them and several host threads drive engines at once: one locked (kernel, device) set.
#include <mutex>
#include <set>
#include <utility>
inline cudaError_t tnsu_opt_in_smem(const void *kernel) {
  static std::mutex mu;
  static std::set<std::pair<const void *, int>> done;
  int dev = 0;
  cudaError_t s = cudaGetDevice(&dev);
  if (s != cudaSuccess) return s;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({kernel, dev})) return cudaSuccess;
  s = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TNSU_SMEM_OPT_IN);
  if (s == cudaSuccess) done.insert({kernel, dev});
  return s;
}

#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__

struct cplx {
  double x, y;
};
HD cplx mkc(double x, double y) {
  cplx r;
  r.x = x;
  r.y = y;
  return r;
}
HD cplx operator+(cplx a, cplx b) { return mkc(a.x + b.x, a.y + b.y); }
HD cplx operator-(cplx a, cplx b) { return mkc(a.x - b.x, a.y - b.y); }
HD cplx operator-(cplx a) { return mkc(-a.x, -a.y); }
HD cplx operator*(cplx a, cplx b) { return mkc(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
HD cplx operator*(cplx a, double b) { return mkc(a.x * b, a.y * b); }
HD cplx operator*(double b, cplx a) { return mkc(a.x * b, a.y * b); }
HD cplx &operator+=(cplx &a, cplx b) {
  a.x += b.x;
  a.y += b.y;
  return a;
}
HD cplx &operator-=(cplx &a, cplx b) {
  a.x -= b.x;
  a.y -= b.y;
  return a;
}

HD double cj(double a) { return a; }
HD cplx cj(cplx a) { return mkc(a.x, -a.y); }
HD double re(double a) { return a; }
HD double re(cplx a) { return a.x; }
HD double abs2(double a) { return a * a; }
HD double abs2(cplx a) { return a.x * a.x + a.y * a.y; }

template <typename S> HD S from_real(double x);
template <> HD double from_real<double>(double x) { return x; }
template <> HD cplx from_real<cplx>(double x) { return mkc(x, 0.0); }
template <typename S> HD S from_c128(double xr, double xi);
template <> HD double from_c128<double>(double xr, double) { return xr; }
template <> HD cplx from_c128<cplx>(double xr, double xi) { return mkc(xr, xi); }
HD cplx to_c128(double a) { return mkc(a, 0.0); }
HD cplx to_c128(cplx a) { return a; }

// acc += a*b with fused multiply-adds
DEV void fma_acc(double &acc, double a, double b) { acc = fma(a, b, acc); }
DEV void fma_acc(cplx &acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc -= a*b
DEV void fms_acc(double &acc, double a, double b) { acc = fma(-a, b, acc); }
DEV void fms_acc(cplx &acc, cplx a, cplx b) {
  acc.x = fma(-a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}

DEV double shfl_xor(double v, int m, unsigned mask = 0xffffffffu) { return __shfl_xor_sync(mask, v, m); }
DEV cplx shfl_xor(cplx v, int m, unsigned mask = 0xffffffffu) {
  return mkc(__shfl_xor_sync(mask, v.x, m), __shfl_xor_sync(mask, v.y, m));
}
DEV double shfl_idx(double v, int l) { return __shfl_sync(0xffffffffu, v, l); }
DEV cplx shfl_idx(cplx v, int l) { return mkc(__shfl_sync(0xffffffffu, v.x, l), __shfl_sync(0xffffffffu, v.y, l)); }

// Named barrier over a sub-group of the CTA (id 1..15; id 0 is __syncthreads).
DEV void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---------------------------------------------------------------------------------------------
// Edge descriptors, built on the host by the scheduler (engine.cu) from the structure matrix and
// the current bond dimensions.  One per edge; rebuilt only when a bond dimension changes.
// ---------------------------------------------------------------------------------------------
struct SideDesc {
  void *base;               // device pointer of tensor storage: point b at base + b*point_stride scalars
  void *ybase;              // scratch of the same geometry: the Householder reflectors Y^H (K x N) between K1 and K3
  long long point_stride;   // capacity in scalars between consecutive points
  int tensor;               // tensor index
  int nlegs;                // z
  int bond_leg;             // 0-based leg (virtual axis) that sits on this edge
  int M, N, K;              // bond matrix rows d*Dk, columns prod(other dims), K = min(M,N)
  int Mout;                 // d*Dnew
  int dims[TNSU_MAX_LEGS];      // leg dimensions before the update
  int leg_edge[TNSU_MAX_LEGS];  // edge id on each leg
  long long in_stride[TNSU_MAX_LEGS];   // element strides of the legs, shape before the update
  long long out_stride[TNSU_MAX_LEGS];  // same, shape after the update (bond dim -> Dnew)
  long long in_stride_phys, out_stride_phys;
};

struct EdgeDesc {
  SideDesc s[2];
  int edge;      // edge id
  int d;         // physical dimension
  int Dk, Dnew;  // bond dimension before / after
  int ni, nj;    // theta is (K_i*d) x (d*K_j)
};

// one site tensor of the network (site RDM kernel, on-chip run kernel)
struct SiteDesc {
  void *base;
  long long point_stride;
  int nlegs;
  int dims[TNSU_MAX_LEGS];
  int leg_edge[TNSU_MAX_LEGS];
  long long nvirt;  // prod dims
};

