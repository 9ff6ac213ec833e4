// Dependent-chain latencies on sm_100a (one warp): nvcc -gencode arch=compute_100a,code=sm_100a -O3 latency.cu -o latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
  __shared__ double sh[256];
  sh[threadIdx.x] = a + threadIdx.x * 1e-3; __syncthreads();
  double x = a + threadIdx.x; float xf = (float)x; long long t0, t1; int idx = threadIdx.x & 31;
  t0 = clock64(); for (int i = 0; i < n; ++i) x = fma(x, a, b); t1 = clock64(); cyc[0] = t1 - t0;           // DFMA chain
  t0 = clock64(); for (int i = 0; i < n; ++i) x = x * a; t1 = clock64(); cyc[1] = t1 - t0;                   // DMUL chain
  t0 = clock64(); for (int i = 0; i < n; ++i) x = x + b; t1 = clock64(); cyc[2] = t1 - t0;                   // DADD chain
  t0 = clock64(); for (int i = 0; i < n; ++i) xf = fmaf(xf, 0.999f, 0.001f); t1 = clock64(); cyc[3] = t1 - t0; // FFMA chain
  t0 = clock64(); for (int i = 0; i < n; ++i) x = rsqrt(x + 2.0); t1 = clock64(); cyc[4] = t1 - t0;          // rsqrt(double) chain (+1 dadd)
  t0 = clock64(); for (int i = 0; i < n; ++i) x = 1.0 / (x + 2.0); t1 = clock64(); cyc[5] = t1 - t0;         // ddiv chain (+1 dadd)
  t0 = clock64(); for (int i = 0; i < n; ++i) x = sqrt(x + 2.0); t1 = clock64(); cyc[6] = t1 - t0;           // dsqrt chain
  t0 = clock64(); for (int i = 0; i < n; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31); t1 = clock64(); cyc[7] = t1 - t0;   // shfl double chain
  t0 = clock64(); for (int i = 0; i < n; ++i) { idx = (int)sh[idx] & 31; } t1 = clock64(); cyc[8] = t1 - t0; // LDS.64 dependent chain (+cvt)
  t0 = clock64(); for (int i = 0; i < n; ++i) { x = fma(sh[(i + idx) & 255], x, b); } t1 = clock64(); cyc[9] = t1 - t0;         // LDS (independent addr) + DFMA chain
  t0 = clock64(); for (int i = 0; i < n; ++i) { __syncthreads(); } t1 = clock64(); cyc[10] = t1 - t0;       // barrier
  out[threadIdx.x] = x + xf + idx;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16);
  const char* names[] = {"DFMA", "DMUL", "DADD", "FFMA", "rsqrt(f64)+dadd", "1/x(f64)+dadd", "sqrt(f64)+dadd", "shfl f64", "LDS.64 dependent (+cvt)", "LDS+DFMA chain", "__syncthreads"};
  for (int threads : {32, 64}) {
    const int n = 2000;
    k<<<1, threads>>>(out, cyc, 0.999, 0.001, n); k<<<1, threads>>>(out, cyc, 0.999, 0.001, n);
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads per CTA = %d\n", threads);
    for (int i = 0; i < 11; ++i) printf("  %-26s %7.1f cycles/op\n", names[i], (double)h[i] / n);
  }
  return 0;
}
