// microbench.cu — B200 FP64 / shuffle / shared-memory issue rates that bound the Simple-Update kernels.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu ; ./microbench
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

__global__ void k_dfma(double *out, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void k_shfl(double *out) {
  int x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 3)) + 1;
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345) out[0] = s;
}

// DFMA + SHFL interleaved: do they overlap?
__global__ void k_mix(double *out, double a, double b) {
  double x[8];
  int y[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) y[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = __shfl_xor_sync(0xffffffffu, y[i], 1 + i) + 1;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += y[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void k_lds128(double *out) {
  __shared__ double2 buf[1024];
  buf[threadIdx.x] = make_double2(threadIdx.x, 1.0);
  __syncthreads();
  double2 acc = make_double2(0, 0);
  int idx = threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double2 v = buf[(idx + i * 32) & 1023];
      acc.x += v.x;
      acc.y += v.y;
    }
    idx = (idx + 1) & 1023;
  }
  if (acc.x == 12345.678) out[0] = acc.x + acc.y;
}

// dependent-chain latencies (1 warp)
__global__ void k_lat(long long *out, double a, double b) {
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 1024; ++it) x = fma(x, a, b);
  long long t1 = clock64();
  double y = x;
#pragma unroll 1
  for (int it = 0; it < 1024; ++it) y = __shfl_xor_sync(0xffffffffu, y, 1) + b;
  long long t2 = clock64();
  double z = y;
#pragma unroll 1
  for (int it = 0; it < 1024; ++it) z = rsqrt(z + 2.0);
  long long t3 = clock64();
  double w = z;
#pragma unroll 1
  for (int it = 0; it < 1024; ++it) w = sqrt(w + 2.0);
  long long t4 = clock64();
  double u = w;
#pragma unroll 1
  for (int it = 0; it < 1024; ++it) u = 1.0 / (u + 2.0);
  long long t5 = clock64();
  if (threadIdx.x == 0) {
    out[0] = t1 - t0;
    out[1] = t2 - t1;
    out[2] = t3 - t2;
    out[3] = t4 - t3;
    out[4] = t5 - t4;
    out[5] = (long long)u;
  }
}

// DMMA m8n8k4
__global__ void k_dmma(double *out) {
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double *out;
  cudaMalloc(&out, 64);
  long long *lat;
  cudaMalloc(&lat, 64);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("SMs %d clock %d kHz\n", sms, clk_khz);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int thr = warps * 32, grid = sms * 2048 / thr / 2;
    grid = sms * (warps >= 32 ? 1 : 2);
    float ms = time_it([&] { k_dfma<<<grid, thr>>>(out, 1.0000001, 1e-9); });
    double fl = (double)grid * thr * ITER * 8 * 2;
    printf("DFMA  warps/CTA %2d CTAs %4d: %.3f ms  %.2f TFLOP/s  (%.1f DFMA lanes/clk/SM at %d kHz)\n", warps, grid, ms,
           fl / ms / 1e9, fl / 2 / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz);
    ms = time_it([&] { k_shfl<<<grid, thr>>>(out); });
    double sh = (double)grid * thr * ITER * 8;
    printf("SHFL  warps/CTA %2d: %.3f ms  %.1f SHFL lanes/clk/SM\n", warps, ms, sh / (ms * 1e-3) / sms / (clk_khz * 1e3));
    ms = time_it([&] { k_mix<<<grid, thr>>>(out, 1.0000001, 1e-9); });
    printf("MIX   warps/CTA %2d: %.3f ms  (8 DFMA + 4 SHFL per iter: %.1f DFMA lanes/clk/SM, %.1f SHFL lanes/clk/SM)\n", warps, ms,
           (double)grid * thr * ITER * 8 / (ms * 1e-3) / sms / (clk_khz * 1e3),
           (double)grid * thr * ITER * 4 / (ms * 1e-3) / sms / (clk_khz * 1e3));
    ms = time_it([&] { k_lds128<<<grid, thr>>>(out); });
    printf("LDS128 warps/CTA %2d: %.3f ms  %.1f B/clk/SM\n", warps, ms,
           (double)grid * thr * ITER * 8 * 16 / (ms * 1e-3) / sms / (clk_khz * 1e3));
    ms = time_it([&] { k_dmma<<<grid, thr>>>(out); });
    double dm = (double)grid * warps * ITER * 4 * (8 * 8 * 4 * 2);
    printf("DMMA  warps/CTA %2d: %.3f ms  %.2f TFLOP/s\n", warps, ms, dm / ms / 1e9);
  }
  k_lat<<<1, 32>>>(lat, 1.0000001, 1e-9);
  long long h[6];
  cudaMemcpy(h, lat, sizeof h, cudaMemcpyDeviceToHost);
  printf("latency cycles/op: DFMA %.1f  SHFL64+DADD %.1f  rsqrt(f64) %.1f  sqrt(f64) %.1f  div(f64) %.1f\n", h[0] / 1024.0,
         h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, h[4] / 1024.0);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
