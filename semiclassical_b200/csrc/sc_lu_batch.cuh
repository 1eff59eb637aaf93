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
  if (dr > 32 && !getenv("SC_LU_DFMA")) {
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
