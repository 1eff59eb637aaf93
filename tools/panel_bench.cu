// latency of the in-warp 4-column panel factorisation (sc_lu.cuh: lu_panel) and of one panel application
#include <cstdio>
#include <cuda_runtime.h>
#include "../semiclassical_b200/csrc/sc_lu.cuh"
using namespace sc;
__global__ void k(const double2 *A, double2 *out, long long *cyc, int reps, int nwarps_active) {
  __shared__ LuPanel pan[2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double2 lo0[4], hi0[4];
  for (int c = 0; c < 4; ++c) { lo0[c] = A[c * 64 + lane]; hi0[c] = A[c * 64 + lane + 32]; }
  double2 acc = make_double2(0, 0);
  __syncthreads();
  if (w < nwarps_active) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      double2 lo[4], hi[4];
      for (int c = 0; c < 4; ++c) { lo[c] = lo0[c]; hi[c] = hi0[c]; lo[c].x += 1e-9 * r; }
      unsigned long long done = 0;
      lu_panel(lo, hi, 4, done, true, &pan[w & 1], lane);
      acc.x += lo[3].x + hi[2].y;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / reps;
    __syncwarp();
    // one panel application to a 4-column block
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      double2 lo[4], hi[4];
      for (int c = 0; c < 4; ++c) { lo[c] = lo0[c]; hi[c] = hi0[c]; lo[c].x += 1e-9 * r; }
      const LuPanel *P = &pan[w & 1];
      for (int c = 0; c < 4; ++c) {
        const int p = P->p[c];
        const double2 flo = P->f[c][lane], fhi = P->f[c][lane + 32];
        for (int j = 0; j < 4; ++j) lu_rank1(lo[j], hi[j], p, flo, fhi, true);
      }
      acc.x += lo[3].x + hi[2].y;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0) / reps;
  }
  out[threadIdx.x] = acc;
}
int main() {
  double2 h[256];
  srand(3);
  for (int i = 0; i < 256; ++i) { h[i].x = rand() / (double)RAND_MAX - 0.5; h[i].y = rand() / (double)RAND_MAX - 0.5; }
  double2 *dA, *dout; long long *dc, hc[2];
  cudaMalloc(&dA, sizeof(h)); cudaMalloc(&dout, sizeof(double2) * 1024); cudaMalloc(&dc, 16);
  cudaMemcpy(dA, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int nw : {1, 2, 4, 8}) {
    k<<<1, 256>>>(dA, dout, dc, 200, nw);
    cudaDeviceSynchronize();
    cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
    printf("active warps %d: lu_panel %lld cycles, apply(4 ranks x 4 cols) %lld cycles  %s\n", nw, hc[0], hc[1], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
