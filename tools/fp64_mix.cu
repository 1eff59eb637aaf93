// Do DMMA (FP64 tensor) and DFMA (FP64 vector) share one pipe on B200?  Warps [0, NT) run register-resident
// DMMA.8x8x4 chains, warps [NT, NT+NV) run DFMA chains; each class is timed alone and together.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(1024, 1) k(int nt, int nv, int iters_t, int iters_v, double *out, long long *cyc) {
  const int w = threadIdx.x >> 5;
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  __syncthreads();
  const long long t0 = clock64();
  if (w < nt) {
    for (int it = 0; it < iters_t; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
  } else if (w < nt + nv) {
    for (int it = 0; it < iters_v; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + w] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8 * 148 * 32);
  long long h[32];
  const int IT = 20000, IV = 20000;
  int cfgs[][2] = {{4, 0}, {0, 4}, {4, 4}, {8, 0}, {0, 8}, {8, 8}, {12, 0}, {12, 4}, {12, 8}, {0, 12}};
  for (auto &cf : cfgs) {
    const int nt = cf[0], nv = cf[1];
    k<<<148, 32 * (nt + nv)>>>(nt, nv, IT, IV, out, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mt = 0, mv = 0;
    for (int w = 0; w < nt; ++w) mt = h[w] > mt ? h[w] : mt;
    for (int w = nt; w < nt + nv; ++w) mv = h[w] > mv ? h[w] : mv;
    // flop per cycle per SM
    const double ft = nt ? (double)nt * IT * 8 * 512 / mt : 0, fv = nv ? (double)nv * IV * 16 * 64 / mv : 0;
    printf("tensor warps %2d, vector warps %2d: DMMA %.1f flop/clk/SM (%lld cyc), DFMA %.1f flop/clk/SM (%lld cyc)  %s\n", nt, nv, ft, mt, fv, mv, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
