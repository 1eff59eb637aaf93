// instruction latency probes on sm_100a (single warp / single CTA), clock64 around dependent chains
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, double x0, int n) {
  __shared__ __align__(16) double sm[1024];
  __shared__ unsigned su[64];
  const int t = threadIdx.x;
  sm[t % 1024] = x0 + t; su[t % 64] = t;
  __syncthreads();
  double x = x0;
  long long t0, t1;
  // 0: dependent DFMA
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, 1.0000001, 1e-9);
  t1 = clock64(); if (t == 0) cyc[0] = (t1 - t0);
  // 1: dependent DMUL+DADD pairs
  t0 = clock64();
  for (int i = 0; i < n; ++i) { x = x * 1.0000001; x = x + 1e-9; }
  t1 = clock64(); if (t == 0) cyc[1] = (t1 - t0) / 2;
  // 2: dependent reciprocal 1/x
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = 1.0 / (x + 0.5);
  t1 = clock64(); if (t == 0) cyc[2] = (t1 - t0);
  // 3: dependent LDS.64 pointer chase
  int idx = t & 31;
  t0 = clock64();
  for (int i = 0; i < n; ++i) idx = (int)su[idx & 63] & 63;
  t1 = clock64(); if (t == 0) cyc[3] = (t1 - t0);
  x += idx;
  // 4: redux.sync max
  unsigned u = t;
  t0 = clock64();
  for (int i = 0; i < n; ++i) u = __reduce_max_sync(0xffffffffu, u + i);
  t1 = clock64(); if (t == 0) cyc[4] = (t1 - t0);
  x += u;
  // 5: __syncthreads
  t0 = clock64();
  for (int i = 0; i < n; ++i) __syncthreads();
  t1 = clock64(); if (t == 0) cyc[5] = (t1 - t0);
  // 6: STS then __syncthreads then LDS (store->barrier->load round trip)
  t0 = clock64();
  for (int i = 0; i < n; ++i) { sm[(t + i) & 1023] = x; __syncthreads(); x += sm[(t + i + 33) & 1023]; }
  t1 = clock64(); if (t == 0) cyc[6] = (t1 - t0);
  // 7: shfl double (2 shuffles)
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  t1 = clock64(); if (t == 0) cyc[7] = (t1 - t0);
  // 8: sqrt
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = sqrt(x + 2.0);
  t1 = clock64(); if (t == 0) cyc[8] = (t1 - t0);
  // 9: exp
  t0 = clock64();
  for (int i = 0; i < n; ++i) x = exp(-x * 1e-3);
  t1 = clock64(); if (t == 0) cyc[9] = (t1 - t0);
  // 10: dependent DMMA chain (same accumulator)
  {
    double c0 = x, c1 = x * 0.5, a = 1e-3, b = 1e-3;
    t0 = clock64();
    for (int i = 0; i < n; ++i) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    t1 = clock64(); if (t == 0) cyc[10] = (t1 - t0);
    x += c0 + c1;
  }
  // 11: 8 independent DMMA accumulators per iteration (single-warp issue rate)
  {
    double c[8][2]; for (int k = 0; k < 8; ++k) { c[k][0] = x + k; c[k][1] = x - k; }
    double a = 1e-3, b = 1e-3;
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    t1 = clock64(); if (t == 0) cyc[11] = (t1 - t0) / 8;
    for (int k = 0; k < 8; ++k) x += c[k][0] + c[k][1];
  }
  // 12: rcp.approx.ftz.f64 + 2 Newton steps, dependent
  t0 = clock64();
  for (int i = 0; i < n; ++i) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x + 0.5)); r = r * fma(-(x + 0.5), r, 2.0); r = r * fma(-(x + 0.5), r, 2.0); x = r; }
  t1 = clock64(); if (t == 0) cyc[12] = (t1 - t0);
  // 13: shfl 32-bit dependent
  { int v = t; t0 = clock64(); for (int i = 0; i < n; ++i) v = __shfl_sync(0xffffffffu, v + 1, (v + i) & 31); t1 = clock64(); if (t == 0) cyc[13] = (t1 - t0); x += v; }
  // 14: STS.128 -> __syncwarp -> LDS.128 round trip
  { double2 *s2 = reinterpret_cast<double2 *>(sm); t0 = clock64(); for (int i = 0; i < n; ++i) { s2[(t + i) & 255] = make_double2(x, x); __syncwarp(); x += s2[(t + i + 5) & 255].x; __syncwarp(); } t1 = clock64(); if (t == 0) cyc[14] = (t1 - t0); }
  out[blockIdx.x * blockDim.x + t] = x;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16);
  const char* names[] = {"DFMA dep", "DMUL/DADD dep", "1.0/x dep", "LDS.32 chase", "redux.max", "__syncthreads", "STS+BAR+LDS", "shfl double+add", "sqrt", "exp", "DMMA dep", "DMMA indep x8 (per DMMA)", "rcp.approx+2NR", "shfl32 dep", "STS128+syncwarp+LDS128"};
  for (int threads : {32, 128, 384, 512}) {
    const int n = 1000;
    k_lat<<<1, threads>>>(out, cyc, 1.5, n);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 8 * 16, cudaMemcpyDeviceToHost);
    printf("threads=%d:", threads);
    for (int i = 0; i < 15; ++i) printf("  %s=%.1f", names[i], h[i] / (double)n);
    printf("\n");
  }
  return 0;
}
