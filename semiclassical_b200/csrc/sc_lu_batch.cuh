// sc_lu_batch.cuh -- throughput kernel around lu_det_blk: determinants of nmat independent dr x dr complex matrices.
#pragma once
#include <cstdlib>

#include "sc_lu.cuh"
#include "sc_lu_mma.cuh"

namespace sc {

// mats: nmat matrices, row-major [a][b] with ld = dr.  The LU works on the transpose (det A^T = det A): LU row <- b
// (lanes: coalesced 16-byte loads), LU column <- a.
template <int NW, int NBLK>
__global__ void __launch_bounds__(32 * NW, (NW <= 8 ? 2 : 1))
k_lu_batch(const double2 *__restrict__ mats, int dr, int nmat, double2 *__restrict__ det_out) {
  __shared__ LuPanel sh[2];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (int mat = blockIdx.x; mat < nmat; mat += gridDim.x) {
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
    __syncthreads();   // panels of the previous matrix are no longer read
    const double2 det = lu_det_blk<NW, NBLK, 0>(lo, hi, dr, sh, w, lane);
    if (t == 0) det_out[mat] = det;
  }
}

// left-looking dataflow variant (lu_det_left): 4 warps per matrix, 66 KB of shared memory, three matrices per SM
template <int NW>
__global__ void __launch_bounds__(32 * NW, 3)
k_lu_left(const double2 *__restrict__ mats, int dr, int nmat, double2 *__restrict__ det_out) {
  extern __shared__ __align__(16) unsigned char lu_smem[];
  LuFlow *sh = reinterpret_cast<LuFlow *>(lu_smem);
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int nblocks = (dr + 3) >> 2;
  flow_bar_init(sh, t);
  int base = 0;
  unsigned parity = 0;
  for (int mat = blockIdx.x; mat < nmat; mat += gridDim.x) {
    __syncthreads();   // barriers initialised / panels of the previous matrix no longer read
    const double2 det = lu_det_left<NW>(mats + (size_t)mat * dr * dr, dr, dr, sh, base, parity, w, lane);
    base += nblocks;
    parity ^= 1u;
    if (t == 0) det_out[mat] = det;
  }
}

// dr > 64 (harmonic molecules with more than 21 atoms): one CTA per matrix, the matrix in shared memory, Gaussian elimination
// with implicit partial pivoting (lu_det, sc_device.cuh).  Correct for any dr whose matrix fits (dr <= 118); the tuned
// kernels above hold rows in 64-bit masks / 64-row register tiles and stop at dr = 64.
__global__ void __launch_bounds__(256)
k_lu_big(const double2 *__restrict__ mats, int dr, int nmat, double2 *__restrict__ det_out) {
  extern __shared__ __align__(16) unsigned char lub_smem[];
  double2 *Cm = reinterpret_cast<double2 *>(lub_smem);
  double2 *pivbuf = Cm + (size_t)dr * dr;
  int *ibuf = reinterpret_cast<int *>(pivbuf + 2);
  const int t = threadIdx.x;
  for (int mat = blockIdx.x; mat < nmat; mat += gridDim.x) {
    const double2 *A = mats + (size_t)mat * dr * dr;
    __syncthreads();
    for (int i = t; i < dr * dr; i += 256) Cm[i] = A[i];
    __syncthreads();
    const double2 det = lu_det<256>(Cm, dr, ibuf, pivbuf, t, 0);
    if (t == 0) det_out[mat] = det;
  }
}

// dr <= 32: ONE WARP per matrix, lane c holds column c in registers.  The owner of column k finds the pivot among its rows
// >= k (no cross-lane search), the row exchange is a select chain on every lane's own column, the multipliers of column k
// reach the other lanes by shuffles; no shared memory, no barriers, ~20 matrices in flight per SM.  DRM >= dr is the
// compile-time size: rows / columns beyond dr are padded with the identity, which leaves the determinant unchanged.
// (k_lu_batch gives 4 warps to one 18 x 18 matrix and keeps two matrices per SM in flight: 39 % of its samples are
// barrier stalls, profiles/ncu_r02_k_lu_batch_d24.txt.)
// elimination step K as a template recursion: K is a compile-time constant on every path (a `#pragma unroll` over 20-28
// steps is not always honoured, and a runtime K puts the column into local memory).  LPM = lanes per matrix: 32, or 16 with
// two matrices per warp.  The multipliers of column K go through a double-buffered shared-memory row (one 128-bit store
// by the owner, one broadcast load by everybody) instead of four 32-bit shuffles and their register moves per element.
template <int K, int DRM, int LPM>
struct LuWarpStep {
  static __device__ __forceinline__ void run(double2 (&c)[DRM], double2 &det, int &swaps, double2 *buf, int ml) {
    constexpr unsigned FULL = 0xffffffffu;
    int p = K;
    double best = c[K].x * c[K].x + c[K].y * c[K].y;
#pragma unroll
    for (int i = K + 1; i < DRM; ++i) {
      const double m = c[i].x * c[i].x + c[i].y * c[i].y;
      if (m > best) { best = m; p = i; }
    }
    p = __shfl_sync(FULL, p, K, LPM);
    // rows K <-> p of the own column
    const double2 ck = c[K];
    double2 cp = ck;
    if (p != K) {                                        // uniform over the lanes of a matrix
#pragma unroll
      for (int i = K + 1; i < DRM; ++i)
        if (p == i) { cp = c[i]; c[i] = ck; }
      c[K] = cp;
      ++swaps;
    }
    __syncwarp();
    const double2 pv = make_double2(__shfl_sync(FULL, cp.x, K, LPM), __shfl_sync(FULL, cp.y, K, LPM));
    det = cmul(det, pv);
    if (K + 1 < DRM) {
      const double2 ip = cinv(pv);
      double2 *row = buf + (K & 1) * DRM;
      if (ml == K) {
#pragma unroll
        for (int i = K + 1; i < DRM; ++i) row[i] = cmul(c[i], ip);
      }
      __syncwarp();
#pragma unroll
      for (int i = K + 1; i < DRM; ++i) {
        const double2 f = row[i];
        c[i].x = fma(-f.x, cp.x, fma(f.y, cp.y, c[i].x));
        c[i].y = fma(-f.x, cp.y, fma(-f.y, cp.x, c[i].y));
      }
    }
    LuWarpStep<K + 1, DRM, LPM>::run(c, det, swaps, buf, ml);
  }
};
template <int DRM, int LPM>
struct LuWarpStep<DRM, DRM, LPM> {
  static __device__ __forceinline__ void run(double2 (&)[DRM], double2 &, int &, double2 *, int) {}
};

template <int DRM>
__global__ void __launch_bounds__(128, (DRM <= 16 ? 4 : DRM <= 24 ? 3 : 2))
k_lu_warp(const double2 *__restrict__ mats, int dr, int nmat, double2 *__restrict__ det_out) {
  constexpr int LPM = DRM <= 16 ? 16 : 32, MPW = 32 / LPM;       // lanes per matrix, matrices per warp
  __shared__ double2 lub[4 * MPW][2 * DRM];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ml = lane & (LPM - 1), half = lane / LPM;
  const int wg = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nw = (int)((gridDim.x * blockDim.x) >> 5);
  double2 *buf = lub[warp * MPW + half];
  for (int m0 = wg * MPW; m0 < nmat; m0 += nw * MPW) {
    const int mat = m0 + half;
    const bool live = mat < nmat;
    const double2 *A = mats + (size_t)(live ? mat : m0) * dr * dr;
    double2 c[DRM];
#pragma unroll
    for (int i = 0; i < DRM; ++i)
      c[i] = (i < dr && ml < dr) ? A[(size_t)i * dr + ml] : make_double2(i == ml ? 1.0 : 0.0, 0.0);
    double2 det = make_double2(1.0, 0.0);
    int swaps = 0;
    LuWarpStep<0, DRM, LPM>::run(c, det, swaps, buf, ml);
    if (swaps & 1) { det.x = -det.x; det.y = -det.y; }
    if (ml == 0 && live) det_out[mat] = det;
    __syncwarp();
  }
}

template <int DRM>
static cudaError_t launch_lu_warp(const double2 *mats, int dr, int nmat, double2 *det_out, int sm_count, cudaStream_t st) {
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaError_t ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lu_warp<DRM>, 128, 0);
    if (ce != cudaSuccess) return ce;
    if (per_sm < 1) per_sm = 1;
  }
  const int per_cta = 4 * (DRM <= 16 ? 2 : 1);
  int grid = sm_count * per_sm;
  if (grid > (nmat + per_cta - 1) / per_cta) grid = (nmat + per_cta - 1) / per_cta;
  k_lu_warp<DRM><<<grid, 128, 0, st>>>(mats, dr, nmat, det_out);
  return cudaGetLastError();
}

// NW x NBLK x 4 >= dr for the tuned kernels (dr <= 64)
static cudaError_t launch_lu_batch(const double2 *mats, int dr, int nmat, double2 *det_out, int sm_count, int ctas_per_sm,
                                   cudaStream_t st) {
  if (nmat <= 0) return cudaSuccess;
  if (dr > 64) {
    const size_t smem = sizeof(double2) * ((size_t)dr * dr + 2) + sizeof(int) * (2 * (size_t)dr + 4);
    cudaError_t ce = cudaFuncSetAttribute(k_lu_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return ce;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    int grid = sm_count * per_sm;
    if (grid > nmat) grid = nmat;
    k_lu_big<<<grid, 256, smem, st>>>(mats, dr, nmat, det_out);
    return cudaGetLastError();
  }
  // crossover measured on rotated AS models (d' = d, 148 000 x 16 matrices): d' = 24: k_lu_warp 47 ms / k_lu_mma 51 ms;
  // d' = 26: 86 / 58; d' = 32: 114 / 70.  k_lu_mma wants exactly three CTAs per SM at every size (two: -25 %, four: -35 %)
  int mma_min = 24;
  if (const char *s = getenv("SC_LU_MMA_MIN")) mma_min = atoi(s);
  if (dr > mma_min && !getenv("SC_LU_DFMA")) {
    // trailing updates on the FP64 tensor pipe (sc_lu_mma.cuh); 3 matrices per SM (60 KB of panels each at dr = 60)
    const size_t smem = lum_smem_bytes(dr);
    cudaError_t ce = cudaFuncSetAttribute(k_lu_mma<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return ce;
    ce = cudaFuncSetAttribute(k_lu_mma<4, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (ce != cudaSuccess) return ce;
    int grid = sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 3);
    if (grid > nmat) grid = nmat;
    k_lu_mma<4, 3><<<grid, 128, smem, st>>>(mats, dr, nmat, det_out);
    return cudaGetLastError();
  }
  if (dr > 32 && !getenv("SC_LU_BLK")) {
    cudaError_t ce = cudaFuncSetAttribute(k_lu_left<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LuFlow));
    if (ce != cudaSuccess) return ce;
    int grid = sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 3);
    if (grid > nmat) grid = nmat;
    k_lu_left<4><<<grid, 128, sizeof(LuFlow), st>>>(mats, dr, nmat, det_out);
    return cudaGetLastError();
  }
  if (dr <= 32 && !getenv("SC_LU_BLK")) {
    if (dr <= 8) return launch_lu_warp<8>(mats, dr, nmat, det_out, sm_count, st);
    if (dr <= 12) return launch_lu_warp<12>(mats, dr, nmat, det_out, sm_count, st);
    if (dr <= 16) return launch_lu_warp<16>(mats, dr, nmat, det_out, sm_count, st);
    if (dr <= 20) return launch_lu_warp<20>(mats, dr, nmat, det_out, sm_count, st);
    if (dr <= 24) return launch_lu_warp<24>(mats, dr, nmat, det_out, sm_count, st);
    if (dr <= 28) return launch_lu_warp<28>(mats, dr, nmat, det_out, sm_count, st);
    return launch_lu_warp<32>(mats, dr, nmat, det_out, sm_count, st);
  }
  if (dr > 32) {
    int grid = sm_count * 2;
    if (grid > nmat) grid = nmat;
    k_lu_batch<8, 2><<<grid, 256, 0, st>>>(mats, dr, nmat, det_out);
  } else if (dr > 16) {
    int grid = sm_count * 4;
    if (grid > nmat) grid = nmat;
    k_lu_batch<4, 2><<<grid, 128, 0, st>>>(mats, dr, nmat, det_out);
  } else {
    int grid = sm_count * 8;
    if (grid > nmat) grid = nmat;
    k_lu_batch<2, 2><<<grid, 64, 0, st>>>(mats, dr, nmat, det_out);
  }
  return cudaGetLastError();
}

}  // namespace sc
