// microbenchmark of the in-CTA complex LU determinant (sc_device.cuh) on one SM-resident CTA per SM
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../semiclassical_b200/csrc/sc_device.cuh"
using namespace sc;
#ifndef VARIANT
#define VARIANT lu_det_cta
#endif
template <int TPT, int RC>
__global__ void __launch_bounds__(TPT, 1) k(const double2* A, int dr, int reps, double2* out, long long* cyc) {
  extern __shared__ __align__(16) double sm[];
  double2* Cm = reinterpret_cast<double2*>(sm);
  unsigned* wkey = reinterpret_cast<unsigned*>(sm + 2 * dr * dr + 200 * 1024 / 8 - 2*60*60 - 64);
  double2 det = make_double2(0, 0);
  long long tot = 0;
  for (int r = 0; r < reps; ++r) {
    const int ldc = RC ? (dr | 1) : dr;
    for (int i = threadIdx.x; i < dr * dr; i += TPT) Cm[(i / dr) * ldc + i % dr] = A[i];
    __syncthreads();
    long long t0 = clock64();
    if (RC) det = lu_det_rc<TPT>(Cm, dr, ldc, wkey, threadIdx.x);
    else det = lu_det_cta<TPT>(Cm, dr, wkey, threadIdx.x);
    __syncthreads();
    tot += clock64() - t0;
  }
  if (threadIdx.x == 0) { out[blockIdx.x] = det; cyc[blockIdx.x] = tot / reps; }
}
template <int TPT, int RC> void run(const double2* dA, int dr, double2* dout, long long* dcyc) {
  size_t smem = 200 * 1024;
  cudaFuncSetAttribute(k<TPT, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<TPT, RC><<<148, TPT, smem>>>(dA, dr, 20, dout, dcyc);
  cudaDeviceSynchronize();
  double2 det; long long c;
  cudaMemcpy(&det, dout, sizeof(det), cudaMemcpyDeviceToHost);
  cudaMemcpy(&c, dcyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("rc=%d TPT=%4d dr=%d: %lld cycles per LU (%.0f per column)  det=(%.12e, %.12e)  %s\n", RC, TPT, dr, c, (double)c / dr, det.x, det.y, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int dr = 60;
  double2* hA = (double2*)malloc(sizeof(double2) * dr * dr);
  srand(1);
  for (int i = 0; i < dr * dr; ++i) { hA[i].x = rand() / (double)RAND_MAX - 0.5; hA[i].y = rand() / (double)RAND_MAX - 0.5; }
  for (int i = 0; i < dr; ++i) hA[i * dr + i].x += 2.0;
  double2 *dA, *dout; long long* dcyc;
  cudaMalloc(&dA, sizeof(double2) * dr * dr); cudaMalloc(&dout, sizeof(double2) * 148); cudaMalloc(&dcyc, 8 * 148);
  cudaMemcpy(dA, hA, sizeof(double2) * dr * dr, cudaMemcpyHostToDevice);
  run<384, 0>(dA, dr, dout, dcyc);
  run<128, 1>(dA, dr, dout, dcyc);
  run<256, 1>(dA, dr, dout, dcyc);
  run<384, 1>(dA, dr, dout, dcyc);
  run<512, 1>(dA, dr, dout, dcyc);
  run<640, 1>(dA, dr, dout, dcyc);
  return 0;
}
