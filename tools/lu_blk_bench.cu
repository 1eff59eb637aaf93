// microbenchmark + self-check of the batched blocked register LU (semiclassical_b200/csrc/sc_lu.cuh)
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../semiclassical_b200/csrc/sc_lu.cuh"
#include "../semiclassical_b200/csrc/sc_lu_batch.cuh"
using namespace sc;
typedef std::complex<double> cd;
static cd cpu_det(std::vector<cd> a, int n) {
  cd det = 1.0;
  for (int k = 0; k < n; ++k) {
    int p = k; double best = std::abs(a[k * n + k]);
    for (int i = k + 1; i < n; ++i) if (std::abs(a[i * n + k]) > best) { best = std::abs(a[i * n + k]); p = i; }
    if (p != k) { for (int j = 0; j < n; ++j) std::swap(a[k * n + j], a[p * n + j]); det = -det; }
    det *= a[k * n + k];
    for (int i = k + 1; i < n; ++i) { cd f = a[i * n + k] / a[k * n + k]; for (int j = k + 1; j < n; ++j) a[i * n + j] -= f * a[k * n + j]; }
  }
  return det;
}
template <int NW, int NBLK, int OCC>
__global__ void __launch_bounds__(32 * NW, OCC) k_time(const double2 *__restrict__ mats, int dr, int nmat, double2 *det_out, long long *cyc) {
  __shared__ LuPanel sh[2];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  long long tl = 0, tu = 0; int cnt = 0;
  for (int mat = blockIdx.x; mat < nmat; mat += gridDim.x) {
    const long long t0 = clock64();
    const double2 *A = mats + (size_t)mat * dr * dr;
    double2 lo[NBLK][4], hi[NBLK][4];
#pragma unroll
    for (int s = 0; s < NBLK; ++s)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = 4 * (w + NW * s) + c;
        lo[s][c] = hi[s][c] = make_double2(0.0, 0.0);
        if (col < dr) {
          if (lane < dr) lo[s][c] = A[(size_t)col * dr + lane];
          if (lane + 32 < dr) hi[s][c] = A[(size_t)col * dr + lane + 32];
        }
      }
    __syncthreads();
    const long long t1 = clock64();
    const double2 det = lu_det_blk<NW, NBLK, 0>(lo, hi, dr, sh, w, lane);
    const long long t2 = clock64();
    if (t == 0) det_out[mat] = det;
    tl += t1 - t0; tu += t2 - t1; ++cnt;
  }
  if (t == 0 && blockIdx.x == 0) { cyc[0] = tl / cnt; cyc[1] = tu / cnt; }
}
template <int NW, int NBLK, int OCC>
__global__ void __launch_bounds__(32 * NW, OCC) k_flow(const double2 *__restrict__ mats, int dr, int nmat, double2 *det_out, long long *cyc) {
  extern __shared__ __align__(16) unsigned char fsm[];
  LuFlow *sh = reinterpret_cast<LuFlow *>(fsm);
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int nblocks = (dr + 3) >> 2;
  flow_bar_init(sh, t);
  long long tl = 0, tu = 0; int cnt = 0, base = 0;
  for (int mat = blockIdx.x; mat < nmat; mat += gridDim.x) {
    const long long t0 = clock64();
    const double2 *A = mats + (size_t)mat * dr * dr;
    double2 lo[NBLK][4], hi[NBLK][4];
#pragma unroll
    for (int s = 0; s < NBLK; ++s)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = 4 * (w + NW * s) + c;
        lo[s][c] = hi[s][c] = make_double2(0.0, 0.0);
        if (col < dr) {
          if (lane < dr) lo[s][c] = A[(size_t)col * dr + lane];
          if (lane + 32 < dr) hi[s][c] = A[(size_t)col * dr + lane + 32];
        }
      }
    __syncthreads();
    const long long t1 = clock64();
    const double2 det = lu_det_flow<NW, NBLK>(lo, hi, dr, sh, base, w, lane);
    base += nblocks;
    const long long t2 = clock64();
    if (t == 0) det_out[mat] = det;
    tl += t1 - t0; tu += t2 - t1; ++cnt;
  }
  if (t == 0 && blockIdx.x == 0) { cyc[0] = tl / cnt; cyc[1] = tu / cnt; }
}
template <int NW, int NBLK, int OCC> void timeflow(const double2 *dA, int dr, int nmat, double2 *ddet, int ctas_per_sm, const std::vector<cd> &h) {
  long long *dc, hc[2]; cudaMalloc(&dc, 16);
  cudaFuncSetAttribute(k_flow<NW, NBLK, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LuFlow));
  cudaMemset(ddet, 0, sizeof(double2) * nmat);
  k_flow<NW, NBLK, OCC><<<148 * ctas_per_sm, 32 * NW, sizeof(LuFlow)>>>(dA, dr, nmat, ddet, dc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_flow<NW, NBLK, OCC><<<148 * ctas_per_sm, 32 * NW, sizeof(LuFlow)>>>(dA, dr, nmat, ddet, dc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
  std::vector<cd> det(nmat);
  cudaMemcpy(det.data(), ddet, sizeof(double2) * nmat, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < nmat; m += (nmat / 61 > 0 ? nmat / 61 : 1)) {
    std::vector<cd> a(h.begin() + (size_t)m * dr * dr, h.begin() + (size_t)(m + 1) * dr * dr);
    const cd ref = cpu_det(a, dr);
    const double err = std::abs(det[m] - ref) / std::abs(ref);
    if (err > maxerr) maxerr = err;
  }
  printf("  FLOW NW=%d NBLK=%d occ=%d ctas/SM=%d: load %lld cyc, LU %lld cyc per matrix (CTA 0); %.3f ms -> %.0f SM-cycles/matrix  err %.1e  %s\n", NW, NBLK, OCC, ctas_per_sm, hc[0], hc[1], ms, ms * 1e-3 * 1.965e9 * 148 / nmat, maxerr, cudaGetErrorString(cudaGetLastError()));
  cudaFree(dc);
}
template <int NW, int OCC>
__global__ void __launch_bounds__(32 * NW, OCC) k_left(const double2 *__restrict__ mats, int dr, int nmat, double2 *det_out, long long *cyc) {
  extern __shared__ __align__(16) unsigned char fsm[];
  LuFlow *sh = reinterpret_cast<LuFlow *>(fsm);
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int nblocks = (dr + 3) >> 2;
  flow_bar_init(sh, t);
  long long tu = 0; int cnt = 0, base = 0; unsigned parity = 0;
  for (int mat = blockIdx.x; mat < nmat; mat += gridDim.x) {
    __syncthreads();
    const long long t1 = clock64();
    const double2 det = lu_det_left<NW>(mats + (size_t)mat * dr * dr, dr, dr, sh, base, parity, w, lane);
    base += nblocks; parity ^= 1u;
    const long long t2 = clock64();
    if (t == 0) det_out[mat] = det;
    tu += t2 - t1; ++cnt;
  }
  if (t == 0 && blockIdx.x == 0) { cyc[0] = 0; cyc[1] = tu / cnt; }
}
template <int NW, int OCC> void timeleft(const double2 *dA, int dr, int nmat, double2 *ddet, int ctas_per_sm, const std::vector<cd> &h) {
  long long *dc, hc[2]; cudaMalloc(&dc, 16);
  cudaFuncSetAttribute(k_left<NW, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LuFlow));
  cudaMemset(ddet, 0, sizeof(double2) * nmat);
  k_left<NW, OCC><<<148 * ctas_per_sm, 32 * NW, sizeof(LuFlow)>>>(dA, dr, nmat, ddet, dc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_left<NW, OCC><<<148 * ctas_per_sm, 32 * NW, sizeof(LuFlow)>>>(dA, dr, nmat, ddet, dc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
  std::vector<cd> det(nmat);
  cudaMemcpy(det.data(), ddet, sizeof(double2) * nmat, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < nmat; m += (nmat / 61 > 0 ? nmat / 61 : 1)) {
    std::vector<cd> a(h.begin() + (size_t)m * dr * dr, h.begin() + (size_t)(m + 1) * dr * dr);
    const cd ref = cpu_det(a, dr);
    const double err = std::abs(det[m] - ref) / std::abs(ref);
    if (err > maxerr) maxerr = err;
  }
  printf("  LEFT NW=%d occ=%d ctas/SM=%d: LU %lld cyc per matrix (CTA 0); %.3f ms -> %.0f SM-cycles/matrix  err %.1e  %s\n", NW, OCC, ctas_per_sm, hc[1], ms, ms * 1e-3 * 1.965e9 * 148 / nmat, maxerr, cudaGetErrorString(cudaGetLastError()));
  cudaFree(dc);
}
template <int NW, int NBLK, int OCC> void timeit(const double2 *dA, int dr, int nmat, double2 *ddet, int ctas_per_sm) {
  long long *dc, hc[2]; cudaMalloc(&dc, 16);
  k_time<NW, NBLK, OCC><<<148 * ctas_per_sm, 32 * NW>>>(dA, dr, nmat, ddet, dc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_time<NW, NBLK, OCC><<<148 * ctas_per_sm, 32 * NW>>>(dA, dr, nmat, ddet, dc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
  printf("  NW=%d NBLK=%d occ=%d ctas/SM=%d: load %lld cyc, LU %lld cyc per matrix (CTA 0); %.3f ms -> %.0f SM-cycles/matrix  %s\n", NW, NBLK, OCC, ctas_per_sm, hc[0], hc[1], ms, ms * 1e-3 * 1.965e9 * 148 / nmat, cudaGetErrorString(cudaGetLastError()));
  cudaFree(dc);
}
int main(int argc, char **argv) {
  const int nmat = argc > 1 ? atoi(argv[1]) : 148 * 32;
  const bool prof = argc > 2;
  int drs[] = {60, 51, 45, 33, 24, 17, 62};
  for (int dr : drs) {
    if (prof && dr != 60) continue;
    std::vector<cd> h((size_t)nmat * dr * dr);
    srand(dr);
    for (int m = 0; m < nmat; ++m)
      for (int i = 0; i < dr * dr; ++i) {
        const bool diag_only = (m % 3 == 2);
        const int r = i / dr, c = i % dr;
        cd v(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
        if (diag_only && r != c) v = 0.0;
        if (r == c && m % 3 != 1) v += cd(1.5, 0.3);
        h[(size_t)m * dr * dr + i] = v;
      }
    double2 *dA, *ddet;
    cudaMalloc(&dA, sizeof(double2) * h.size());
    cudaMalloc(&ddet, sizeof(double2) * nmat);
    cudaMemcpy(dA, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch_lu_batch(dA, dr, nmat, ddet, 148, 3, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch_lu_batch(dA, dr, nmat, ddet, 148, 3, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<cd> det(nmat);
    cudaMemcpy(det.data(), ddet, sizeof(double2) * nmat, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < nmat; m += (nmat / 97 > 0 ? nmat / 97 : 1)) {
      std::vector<cd> a(h.begin() + (size_t)m * dr * dr, h.begin() + (size_t)(m + 1) * dr * dr);
      const cd ref = cpu_det(a, dr);
      const double err = std::abs(det[m] - ref) / std::abs(ref);
      if (err > maxerr) maxerr = err;
    }
    printf("dr=%2d nmat=%d: %.3f ms  %.0f SM-cycles/matrix (1965 MHz)  max rel err vs CPU LU = %.2e  %s\n", dr, nmat, ms,
           ms * 1e-3 * 1.965e9 * 148 / nmat, maxerr, cudaGetErrorString(cudaGetLastError()));
    if (dr == 60 && prof) {
      timeflow<8, 2, 2>(dA, dr, nmat, ddet, 2, h);
    } else if (dr == 60) {
      timeit<8, 2, 1>(dA, dr, nmat, ddet, 1);
      timeit<8, 2, 2>(dA, dr, nmat, ddet, 1);
      timeit<8, 2, 2>(dA, dr, nmat, ddet, 2);
      timeit<16, 1, 1>(dA, dr, nmat, ddet, 1);
      timeit<16, 1, 2>(dA, dr, nmat, ddet, 2);
      timeit<4, 4, 1>(dA, dr, nmat, ddet, 1);
      timeit<4, 4, 3>(dA, dr, nmat, ddet, 3);
      timeleft<8, 1>(dA, dr, nmat, ddet, 1, h);
      timeleft<8, 3>(dA, dr, nmat, ddet, 3, h);
      timeleft<4, 3>(dA, dr, nmat, ddet, 3, h);
      timeleft<16, 1>(dA, dr, nmat, ddet, 1, h);
      timeleft<8, 2>(dA, dr, nmat, ddet, 2, h);
      timeleft<4, 4>(dA, dr, nmat, ddet, 3, h);
      timeleft<5, 3>(dA, dr, nmat, ddet, 3, h);
      timeflow<8, 2, 1>(dA, dr, nmat, ddet, 1, h);
      timeflow<8, 2, 2>(dA, dr, nmat, ddet, 2, h);
      timeflow<16, 1, 1>(dA, dr, nmat, ddet, 1, h);
      timeflow<4, 4, 1>(dA, dr, nmat, ddet, 1, h);
      timeflow<4, 4, 3>(dA, dr, nmat, ddet, 3, h);
    }
    cudaFree(dA); cudaFree(ddet);
  }
  return 0;
}
