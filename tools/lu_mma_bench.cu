// self-check + timing of the DMMA left-looking LU (semiclassical_b200/csrc/sc_lu_mma.cuh) against a CPU LU and
// against the DFMA left-looking kernels (sc_lu_batch.cuh)
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../semiclassical_b200/csrc/sc_lu_mma.cuh"
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
template <int OCC, int NW = 4> static float run_mma(const double2 *dA, int dr, int nmat, double2 *ddet, int ctas_per_sm) {
  const size_t smem = lum_smem_bytes(dr);
  cudaFuncSetAttribute(k_lu_mma<NW, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = 148 * ctas_per_sm; if (grid > nmat) grid = nmat;
  k_lu_mma<NW, OCC><<<grid, 32 * NW, smem>>>(dA, dr, nmat, ddet);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_lu_mma<NW, OCC><<<grid, 32 * NW, smem>>>(dA, dr, nmat, ddet);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
int main(int argc, char **argv) {
  const int nmat = argc > 1 ? atoi(argv[1]) : 148 * 48;
  const int only = argc > 2 ? atoi(argv[2]) : 0;
  const int diag = argc > 3 ? atoi(argv[3]) : 0;   // 1: all matrices diagonal (the separable AS model)
  int drs[] = {60, 64, 33, 45, 51, 62, 37};
  for (int dr : drs) {
    if (only && dr != only) continue;
    std::vector<cd> h((size_t)nmat * dr * dr);
    srand(dr);
    for (int m = 0; m < nmat; ++m)
      for (int i = 0; i < dr * dr; ++i) {
        const bool diag_only = diag || (m % 4 == 2);
        const int r = i / dr, c = i % dr;
        cd v(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
        if (m % 4 == 3) v *= 0.05;                      // near-identity (the early-time prefactor matrices)
        if (diag_only && r != c) v = 0.0;
        if (r == c && m % 4 != 1) v += cd(1.5, 0.3);     // m % 4 == 1: fully random, pivoting essential
        h[(size_t)m * dr * dr + i] = v;
      }
    double2 *dA, *ddet;
    cudaMalloc(&dA, sizeof(double2) * h.size());
    cudaMalloc(&ddet, sizeof(double2) * nmat);
    cudaMemcpy(dA, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice);
    auto check = [&](const char *tag, float ms) {
      std::vector<cd> det(nmat);
      cudaMemcpy(det.data(), ddet, sizeof(double2) * nmat, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < nmat; m += (nmat / 203 > 0 ? nmat / 203 : 1)) {
        std::vector<cd> a(h.begin() + (size_t)m * dr * dr, h.begin() + (size_t)(m + 1) * dr * dr);
        const cd ref = cpu_det(a, dr);
        const double err = std::abs(det[m] - ref) / std::abs(ref);
        if (!(err <= maxerr)) maxerr = err;
      }
      printf("dr=%2d %-22s %.3f ms  %6.0f SM-cycles/matrix  max rel err vs CPU LU = %.2e  %s\n", dr, tag, ms,
             ms * 1e-3 * 1.965e9 * 148 / nmat, maxerr, cudaGetErrorString(cudaGetLastError()));
    };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    setenv("SC_LU_DFMA", "1", 1);   // the DFMA left-looking kernel k_lu_left<4> for comparison
    launch_lu_batch(dA, dr, nmat, ddet, 148, 0, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch_lu_batch(dA, dr, nmat, ddet, 148, 0, 0);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    check("DFMA left-looking", ms);
    unsetenv("SC_LU_DFMA");
    cudaMemset(ddet, 0, sizeof(double2) * nmat);
    ms = run_mma<3>(dA, dr, nmat, ddet, 3); check("DMMA occ3", ms);
    cudaMemset(ddet, 0, sizeof(double2) * nmat);
    ms = run_mma<2>(dA, dr, nmat, ddet, 2); check("DMMA occ2", ms);
    cudaMemset(ddet, 0, sizeof(double2) * nmat);
    ms = run_mma<1>(dA, dr, nmat, ddet, 1); check("DMMA occ1", ms);
    // warps per matrix at 3 matrices per SM, and occupancy scaling with 2 warps per matrix (small d': more matrices fit)
    ms = run_mma<3, 2>(dA, dr, nmat, ddet, 3); check("DMMA 2 warps occ3", ms);
    ms = run_mma<3, 1>(dA, dr, nmat, ddet, 3); check("DMMA 1 warp occ3", ms);
    ms = run_mma<3, 8>(dA, dr, nmat, ddet, 3); check("DMMA 8 warps occ3", ms);
    ms = run_mma<1, 2>(dA, dr, nmat, ddet, 1); check("DMMA 2 warps occ1", ms);
    ms = run_mma<2, 2>(dA, dr, nmat, ddet, 2); check("DMMA 2 warps occ2", ms);
    if (lum_smem_bytes(dr) * 4 + 4096 < 227 * 1024) { ms = run_mma<4, 2>(dA, dr, nmat, ddet, 4); check("DMMA 2 warps occ4", ms); }
    if (lum_smem_bytes(dr) * 5 + 5120 < 227 * 1024) { ms = run_mma<5, 2>(dA, dr, nmat, ddet, 5); check("DMMA 2 warps occ5", ms); }
    cudaMemset(ddet, 0, sizeof(double2) * nmat);
    ms = run_mma<3, 2>(dA, dr, nmat, ddet, 3); check("DMMA 2 warps occ3", ms);
    cudaMemset(ddet, 0, sizeof(double2) * nmat);
    ms = run_mma<3, 1>(dA, dr, nmat, ddet, 3); check("DMMA 1 warp occ3", ms);
    cudaMemset(ddet, 0, sizeof(double2) * nmat);
    ms = run_mma<3, 8>(dA, dr, nmat, ddet, 3); check("DMMA 8 warps occ3", ms);
    // occupancy scaling with 2 warps per matrix (small d': more matrices fit)
    ms = run_mma<1, 2>(dA, dr, nmat, ddet, 1); check("DMMA 2 warps occ1", ms);
    ms = run_mma<2, 2>(dA, dr, nmat, ddet, 2); check("DMMA 2 warps occ2", ms);
    if (lum_smem_bytes(dr) * 4 + 4096 < 227 * 1024) { ms = run_mma<4, 2>(dA, dr, nmat, ddet, 4); check("DMMA 2 warps occ4", ms); }
    if (lum_smem_bytes(dr) * 5 + 5120 < 227 * 1024) { ms = run_mma<5, 2>(dA, dr, nmat, ddet, 5); check("DMMA 2 warps occ5", ms); }
    if (lum_smem_bytes(dr) * 6 + 6144 < 227 * 1024) { ms = run_mma<6, 2>(dA, dr, nmat, ddet, 6); check("DMMA 2 warps occ6", ms); }
    cudaFree(dA); cudaFree(ddet);
  }
  return 0;
}
