// latency_bench.cu — dependent-chain latencies (cycles, one warp) of the instruction patterns in the K1 / K2 scalar
// sections.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o latency_bench latency_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
#define T(name, init, ...)                                     \
  {                                                            \
    double x = init;                                           \
    long long t0 = clock64();                                  \
    _Pragma("unroll 1") for (int i = 0; i < N; ++i) { __VA_ARGS__; } \
    long long t1 = clock64();                                  \
    if (threadIdx.x == 0) out[k] = (t1 - t0) / (double)N;      \
    sink += x;                                                 \
    ++k;                                                       \
  }
__global__ void lat(double *out, double a, double b, int zero) {
  __shared__ double sm[64];
  sm[threadIdx.x] = threadIdx.x * 0.5 + 1.0;
  sm[threadIdx.x + 32] = 2.0;
  __syncthreads();
  double sink = 0;
  int k = 0;
  const int lane = threadIdx.x;
  T("loop only", 1.0, x += 0.0 * i)                                         // 0 (int->double + fma)
  T("dfma", 1.0, x = fma(x, a, b))                                          // 1
  T("dadd", 1.0, x = x + b)                                                 // 2
  T("dmul", 1.0, x = x * a)                                                 // 3
  T("dfma x4 chain", 1.0, x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b))   // 4
  T("sqrt", 2.0, x = sqrt(x) + 1.5)                                         // 5
  T("rsqrt", 2.0, x = rsqrt(x) + 1.5)                                       // 6
  T("div", 2.0, x = 1.0 / x + 1.5)                                          // 7
  T("shfl64", 1.0, x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31) + b)    // 8
  T("shfl64 xor", 1.0, x = __shfl_xor_sync(0xffffffffu, x, 1) + b)          // 9
  T("lds", 1.0, x = sm[(int)x & 31] + b)                                    // 10 (cvt + lds + add)
  T("sts+syncwarp+lds", 1.0, sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + b; __syncwarp())  // 11
  T("fsel then dfma", 1.0, x = (x > a) ? fma(x, a, b) : fma(x, b, a))       // 12
  T("divergent if", 1.0, if (lane < zero + 16) x = fma(x, a, b); else x = fma(x, b, a))          // 13
  T("uniform branch", 1.0, if (i >= zero) x = fma(x, a, b))                 // 14
  T("redux+vote", 1.0, { unsigned kk = __reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(x) + lane); x = fma(x, a, (kk & 1) ? b : a); })  // 15
  T("householder scalars", 2.0, {                                           // 16: the K1 chain after the reduction
    double norm2 = fma(x, x, a);
    double norm = sqrt(norm2);
    double uk = fma(1.0, norm, x);
    double tau = 1.0 / (norm * (norm + fabs(x)));
    double t = fma(b, uk, a);
    double w = t * tau;
    x = fma(-w, uk, b) * 1e-3 + 1.0;
  })
  T("jacobi scalars", 2.0, {                                                // 17: the K2 chain after the dots
    double al = x, be = a, ga = b;
    double delta = be - al;
    double g2 = ga * ga;
    double inv_r = rsqrt(fma(delta, delta, 4.0 * g2));
    double c2 = fma(0.5 * fabs(delta), inv_r, 0.5);
    double inv_c = rsqrt(c2);
    double c = c2 * inv_c;
    double s = ga * inv_r * inv_c;
    x = fma(s, al, c * be);
  })
  if (sink == 123.456) out[63] = sink;
}
int main() {
  double *d;
  cudaMalloc(&d, 64 * 8);
  lat<<<1, 32>>>(d, 1.0000001, 1e-9, 0);
  double h[64];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  const char *names[] = {"loop only (i2d + dfma)", "dfma", "dadd", "dmul", "dfma x4 chain (per 4)", "sqrt + dadd", "rsqrt + dadd", "div + dadd", "shfl64 idx + dadd",
                         "shfl64 xor + dadd", "d2i + lds + dadd", "sts + syncwarp + lds + dadd + syncwarp", "dsetp + 2 dfma + fsel", "divergent if/else dfma",
                         "uniform branch + dfma", "redux.max + sel + dfma", "householder scalar chain", "jacobi scalar chain"};
  for (int i = 0; i < 18; ++i) printf("%-42s %7.1f cycles\n", names[i], h[i]);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
