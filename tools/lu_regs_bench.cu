// microbenchmark of the register-resident complex LU determinant (sc_device.cuh: lu_det_regs), one CTA per SM
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../semiclassical_b200/csrc/sc_device.cuh"
using namespace sc;
template <int NW, int MC>
__global__ void __launch_bounds__(32 * NW, 1) k(const double2* A, int dr, int reps, double2* out, long long* cyc) {
  __shared__ LuShared sh;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  double2 det = make_double2(0, 0);
  long long tot = 0;
  for (int r = 0; r < reps; ++r) {
    double2 lo[MC], hi[MC];
#pragma unroll
    for (int m = 0; m < MC; ++m) {
      const int a = warp + NW * m;
      lo[m] = hi[m] = make_double2(0.0, 0.0);
      if (a < dr) {
        if (lane < dr) lo[m] = A[a * dr + lane];
        if (lane + 32 < dr) hi[m] = A[a * dr + lane + 32];
      }
    }
    __syncthreads();
    long long t0 = clock64();
    det = lu_det_regs<NW, MC, 0>(lo, hi, dr, &sh, warp, lane);
    __syncthreads();
    tot += clock64() - t0;
  }
  if (threadIdx.x == 0) { out[blockIdx.x] = det; cyc[blockIdx.x] = tot / reps; }
}
template <int NW, int MC> void run(const double2* dA, int dr, double2* dout, long long* dcyc, const char *what) {
  k<NW, MC><<<148, 32 * NW>>>(dA, dr, 20, dout, dcyc);
  cudaDeviceSynchronize();
  double2 det; long long c;
  cudaMemcpy(&det, dout, sizeof(det), cudaMemcpyDeviceToHost);
  cudaMemcpy(&c, dcyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("%s NW=%2d MC=%d dr=%d: %lld cycles per LU (%.0f per column)  det=(%.12e, %.12e)  %s\n", what, NW, MC, dr, c, (double)c / dr, det.x, det.y, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int dr = 60;
  double2* hA = (double2*)malloc(sizeof(double2) * dr * dr);
  double2 *dA, *dout; long long* dcyc;
  cudaMalloc(&dA, sizeof(double2) * dr * dr); cudaMalloc(&dout, sizeof(double2) * 148); cudaMalloc(&dcyc, 8 * 148);
  for (int pass = 0; pass < 2; ++pass) {
    srand(1);
    for (int i = 0; i < dr * dr; ++i) { hA[i].x = pass ? 0.0 : rand() / (double)RAND_MAX - 0.5; hA[i].y = pass ? 0.0 : rand() / (double)RAND_MAX - 0.5; }
    for (int i = 0; i < dr; ++i) { hA[i * dr + i].x += 2.0 + 0.01 * i; hA[i * dr + i].y += 0.3; }
    cudaMemcpy(dA, hA, sizeof(double2) * dr * dr, cudaMemcpyHostToDevice);
    const char *what = pass ? "diagonal" : "dense   ";
    run<4, 15>(dA, dr, dout, dcyc, what);
    run<8, 8>(dA, dr, dout, dcyc, what);
    run<12, 5>(dA, dr, dout, dcyc, what);
    run<16, 4>(dA, dr, dout, dcyc, what);
  }
  return 0;
}
