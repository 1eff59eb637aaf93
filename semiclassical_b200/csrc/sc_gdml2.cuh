// sc_gdml2.cuh -- sGDML energy / gradient / analytic Hessian, second generation (reference: gdml_predictor.py:140-250
// behind MolecularGDMLPotential.harmonic_approximation, potentials.py:669-699).
//
// Same mathematics as sc_gdml.cuh (sparse descriptor Jacobian, Hessian kernel sum as a rank-M update), restructured after
// its profile: there 50 % of the warp samples sat on the dependent shared-memory loads of the sparse-Jacobian sums
//     XJ[m][x] = sum_p (x - x_m)[p] J[p][x],   AJ[m][x] = sum_p A_m[p] J[p][x]
// (one warp per training point, 51 coordinates on 32 lanes, pair index and Jacobian entry re-fetched per term).  Here
//   * a thread owns ONE coordinate x for the whole geometry: the N-1 signed Jacobian entries of x and their pair indices
//     live in REGISTERS (loaded once per geometry); per training point it needs 2 shared loads per 2 FMA, and it serves
//     4 training points of a tile
//   * training points come in tiles of 16 (cp.async.bulk -> mbarrier, double buffered), two per warp in the distance phase;
//     the per-pair gradient accumulators gx stay in registers across all tiles
//   * the rank-M Hessian update  S += sum_m a_m (x) XJ_m  runs on the FP64 tensor pipe: A fragment = a[m][x], B fragment =
//     XJ[m][y] straight from the tile arrays (leading dimension 72: conflict free), row strip I = warp, 8 x 8 accumulator tiles
//     in registers across all tiles; H = S + S^T + pair blocks at the end
// Envelope: N <= 21 atoms (3N <= 64: one 8-row strip per warp), any number of training points.  Summation order over the
// training points differs from the reference's (tile-wise, DMMA k-steps); the reference's own output moves by ~3e-9
// (absolute) under a permutation of its training set.
#pragma once
#include "sc_gdml.cuh"
#include "sc_mma.cuh"

namespace sc {

constexpr int G2_TM = 16;          // training points per tile
constexpr int G2_THREADS = 256;    // 8 warps
constexpr int G2_LD = 72;          // leading dimension of the XJ / a tiles (== 8 mod 32: conflict-free column fragments)
constexpr int G2_MAXN = 21;        // template NP = N - 1 rounded up (16: N <= 17, 20: N <= 21); PL = pairs per lane
constexpr int G2_SLD = 65;         // leading dimension of the symmetrisation buffer

struct Gdml2Layout {
  int xt, at, xd, xj, aj, xs, jv, gxw, sc, rp, red, total;   // offsets in doubles
};

__host__ __device__ inline Gdml2Layout make_gdml2_layout(int N, int D) {
  Gdml2Layout L;
  const int X = 3 * N;
  int o = 0;
  L.xt = o; o += 2 * G2_TM * D;            // training descriptors, double buffered (later: the 64 x 65 symmetrisation buffer)
  L.at = o; o += 2 * G2_TM * D;            // Jx_alphas rows, double buffered
  if (o < 64 * G2_SLD) o = 64 * G2_SLD;
  L.xd = o; o += G2_TM * D;                // x - x_m of the current tile
  L.xj = o; o += G2_TM * G2_LD;
  L.aj = o; o += G2_TM * G2_LD;
  L.xs = o; o += (D + 1) & ~1;
  L.jv = o; o += 3 * D + (D & 1);
  L.gxw = L.xj;                            // per-warp partial dE/dx_desc (final reduction): the XJ / a tiles are dead then
  L.sc = o; o += 2 * G2_TM;                // ef, w1/2 of the tile's training points
  L.rp = o; o += (X + 1) & ~1;
  L.red = o; o += 32;
  L.total = (o + 1) & ~1;
  return L;
}

template <int NP, int PL>
__global__ void __launch_bounds__(G2_THREADS, (NP <= 16 ? 2 : 1))
k_gdml_eval2(PotDev P, int n, const double *__restrict__ r, double *__restrict__ V, double *__restrict__ grad,
             double *__restrict__ hess, Gdml2Layout L, int hs_ld, size_t hs_stride) {
  extern __shared__ __align__(16) double gsm[];
  __shared__ __align__(8) uint64_t bars[2];
  const int N = P.n_atoms, M = P.n_train, D = P.n_desc, X = 3 * N;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const double q = sqrt(5.0) / P.sig, q2 = q * q, q4_3 = q2 * q2 / 3.0;
  double *xtb = gsm + L.xt, *atb = gsm + L.at, *xd = gsm + L.xd, *XJ = gsm + L.xj, *AJ = gsm + L.aj;
  double *xs = gsm + L.xs, *jv = gsm + L.jv, *gxw = gsm + L.gxw, *sc = gsm + L.sc, *rp = gsm + L.rp, *red = gsm + L.red;
  double *Ssm = gsm + L.xt;
  const int ntiles = (M + G2_TM - 1) / G2_TM;
  const int mtx = (X + 7) >> 3;                      // 8-row strips of the Hessian
  // coordinate role: thread (x, g) serves training points 4 g .. 4 g + 3 of a tile
  const int xrole = t % X, grole = t / X;
  const bool has_x = grole < 4;
  const int xI = xrole / 3, xu = xrole - 3 * xI;
  // upper-triangular atom block (I <= J) owned by this thread in the output phase
  int bI = -1, bJ = -1;
  {
    const int nblk = N * (N + 1) / 2;
    if (t < nblk) {
      int rem = t, I = 0;
      while (rem >= N - I) { rem -= N - I; ++I; }
      bI = I;
      bJ = I + rem;
    }
  }
  if (t == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t phase[2] = {0u, 0u};

  auto issue_tile = [&](int tile) {
    const int b = tile & 1, m0 = tile * G2_TM;
    const int rows = (M - m0 < G2_TM) ? M - m0 : G2_TM;
    if (t == 0) {
      const uint32_t bytes = (uint32_t)(rows * D * sizeof(double));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bars[b], 2 * bytes);
      bulk_g2s(xtb + b * G2_TM * D, P.xs_train + (size_t)m0 * D, bytes, &bars[b]);
      bulk_g2s(atb + b * G2_TM * D, P.jx_alphas + (size_t)m0 * D, bytes, &bars[b]);
    }
  };

  for (int geom = blockIdx.x; geom < n; geom += gridDim.x) {
    // ---- descriptor and the non-zeros of its Jacobian
    for (int x = t; x < X; x += G2_THREADS) rp[x] = r[(size_t)x * n + geom];
    for (int i = t; i < 2 * G2_TM * G2_LD; i += G2_THREADS) XJ[i] = 0.0;      // XJ and a tiles incl. their padding columns
    __syncthreads();                                   // also: the symmetrisation buffer of the previous geometry is free
    issue_tile(0);
    for (int p = t; p < D; p += G2_THREADS) {
      int i = (int)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
      while (i * (i - 1) / 2 > p) --i;
      while ((i + 1) * i / 2 <= p) ++i;
      const int j = p - i * (i - 1) / 2;
      const double dx = rp[3 * i] - rp[3 * j], dy = rp[3 * i + 1] - rp[3 * j + 1], dz = rp[3 * i + 2] - rp[3 * j + 2];
      const double x1 = 1.0 / sqrt(dx * dx + dy * dy + dz * dz), x3 = x1 * x1 * x1;
      xs[p] = x1;
      jv[3 * p] = x3 * dx;       // J[p][3j+u] = +jv, J[p][3i+u] = -jv
      jv[3 * p + 1] = x3 * dy;
      jv[3 * p + 2] = x3 * dz;
    }
    __syncthreads();
    // signed Jacobian entries of this thread's coordinate and their pair indices: registers for the whole geometry
    double sjv[NP];
    unsigned pidx4[NP / 4];                            // pair indices (< 256), four per register
#pragma unroll
    for (int k = 0; k < NP / 4; ++k) pidx4[k] = 0u;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
      sjv[k] = 0.0;
      if (k < N - 1) {
        const int o = k < xI ? k : k + 1;
        const int p = o < xI ? pair_index(xI, o) : pair_index(o, xI);
        pidx4[k >> 2] |= static_cast<unsigned>(p) << (8 * (k & 3));
        sjv[k] = (o < xI ? -1.0 : 1.0) * jv[3 * p + xu];
      }
    }
    double gxa[PL];
#pragma unroll
    for (int k = 0; k < PL; ++k) gxa[k] = 0.0;
    double en = 0.0, sefxa = 0.0;                      // lane 0 of every warp
    double acc[8][2];
#pragma unroll
    for (int J = 0; J < 8; ++J) acc[J][0] = acc[J][1] = 0.0;

    // ---- kernel sum over the training points
    for (int tile = 0; tile < ntiles; ++tile) {
      const int b = tile & 1, m0 = tile * G2_TM;
      if (tile + 1 < ntiles) issue_tile(tile + 1);     // the other buffer was released by the barriers of the previous tile
      mbar_wait(&bars[b], phase[b]);
      phase[b] ^= 1u;
      // phase A: distances, Matern factors, per-pair gradient accumulators; two training points per warp
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int mm = 2 * warp + h, m = m0 + mm;
        const bool live = m < M;
        const double *xt = xtb + (b * G2_TM + mm) * D, *A = atb + (b * G2_TM + mm) * D;
        double *xdw = xd + mm * D;
        double df[PL], av[PL], n2 = 0.0, xa = 0.0;
#pragma unroll
        for (int k = 0; k < PL; ++k) {
          const int p = lane + 32 * k;
          df[k] = av[k] = 0.0;
          if (p < D) {
            if (live) { df[k] = xs[p] - xt[p]; av[k] = A[p]; }
            xdw[p] = df[k];
            n2 = fma(df[k], df[k], n2);
            xa = fma(df[k], av[k], xa);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          n2 += __shfl_xor_sync(0xffffffffu, n2, o);
          xa += __shfl_xor_sync(0xffffffffu, xa, o);
        }
        double ef = 0.0, hw1 = 0.0;
        if (live) {
          const double xn = sqrt(n2);
          ef = q4_3 * exp(-q * xn);
          const double k1 = ef * (1.0 + q * xn) / q2;
          hw1 = 0.5 * ef * xa * q / xn;
          const double efxa = ef * xa;
#pragma unroll
          for (int k = 0; k < PL; ++k) gxa[k] += k1 * av[k] - efxa * df[k];
          en += k1 * xa;
          sefxa += efxa;
        }
        if (lane == 0) { sc[2 * mm] = ef; sc[2 * mm + 1] = hw1; }
      }
      __syncthreads();
      // phase B: XJ_m and a_m = w1/2 XJ_m - ef AJ_m of 4 training points, Jacobian entries from registers
      if (has_x) {
        // k outer, the 4 training points inner: pair index and Jacobian entry are unpacked once per term
        const double *xd0 = xd + (4 * grole) * D, *A0 = atb + (b * G2_TM + 4 * grole) * D;
        double sx[4] = {0.0, 0.0, 0.0, 0.0}, sa[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < NP; ++k)
          if (k < N - 1) {
            const int p = (pidx4[k >> 2] >> (8 * (k & 3))) & 255u;
            const double jk = sjv[k];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              sx[j] = fma(jk, xd0[j * D + p], sx[j]);
              sa[j] = fma(jk, A0[j * D + p], sa[j]);
            }
          }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int mm = 4 * grole + j;
          const bool live = m0 + mm < M;               // rows of the tile buffer beyond M hold stale data
          XJ[mm * G2_LD + xrole] = live ? sx[j] : 0.0;
          AJ[mm * G2_LD + xrole] = live ? sc[2 * mm + 1] * sx[j] - sc[2 * mm] * sa[j] : 0.0;
        }
      }
      __syncthreads();
      // phase C: S += sum_m a_m (x) XJ_m on the tensor pipe; row strip = warp
      if (warp < mtx) {
#pragma unroll
        for (int kk = 0; kk < G2_TM / 4; ++kk) {
          const double af = AJ[(4 * kk + fc) * G2_LD + 8 * warp + fr];
          const double *bp = XJ + (4 * kk + fc) * G2_LD + fr;
#pragma unroll
          for (int J = 0; J < 8; ++J)
            if (J < mtx) dmma884(acc[J][0], acc[J][1], af, bp[8 * J]);
        }
      }
    }
    __syncthreads();                                   // all tiles consumed: the tile buffers become the symmetrisation buffer
    // ---- S to shared memory, reductions over the warps (fixed order)
    if (warp < mtx) {
#pragma unroll
      for (int J = 0; J < 8; ++J)
        if (J < mtx) {
          const int row = 8 * warp + fr, col = 8 * J + 2 * fc;
          Ssm[row * G2_SLD + col] = acc[J][0];
          Ssm[row * G2_SLD + col + 1] = acc[J][1];
        }
    }
#pragma unroll
    for (int k = 0; k < PL; ++k) {
      const int p = lane + 32 * k;
      if (p < D) gxw[warp * D + p] = gxa[k];
    }
    if (lane == 0) { red[warp] = en; red[8 + warp] = sefxa; }
    __syncthreads();
    for (int p = t; p < D; p += G2_THREADS) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += gxw[w * D + p];
      xd[p] = s;                                       // gx (the x - x_m tile is dead)
    }
    double ensum = 0.0, sumefxa = 0.0;
    for (int w = 0; w < 8; ++w) { ensum += red[w]; sumefxa += red[8 + w]; }
    __syncthreads();
    const double *gx = xd;
    if (t == 0) V[geom] = ensum * P.gstd + P.e0 - P.origin;
    if (grad && grole == 0) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < NP; ++k)
        if (k < N - 1) s = fma(sjv[k], gx[(pidx4[k >> 2] >> (8 * (k & 3))) & 255u], s);
      grad[(size_t)xrole * n + geom] = s * P.gstd;
    }
    if (hess && bI >= 0) {
      double hacc[3][3];
#pragma unroll
      for (int u = 0; u < 3; ++u)
#pragma unroll
        for (int v = 0; v < 3; ++v)
          hacc[u][v] = Ssm[(3 * bI + u) * G2_SLD + 3 * bJ + v] + Ssm[(3 * bJ + v) * G2_SLD + 3 * bI + u];
      // pair blocks B_p[u][v] = -sum(ef XA) jv_u jv_v + 3 gx jv_u jv_v / x + delta_uv (-gx x^3):
      // added to the diagonal atom blocks of both atoms, subtracted from the off-diagonal ones (gdml_predictor.py:205-244)
      auto pair_block = [&](int p, double sgn) {
        const double g = gx[p], x1 = xs[p];
        const double c1 = -sumefxa + 3.0 * g / x1, h2 = -g * x1 * x1 * x1;
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
          for (int v = 0; v < 3; ++v) hacc[u][v] += sgn * (c1 * jv[3 * p + u] * jv[3 * p + v] + (u == v ? h2 : 0.0));
      };
      if (bI == bJ) {
        for (int o = 0; o < bI; ++o) pair_block(pair_index(bI, o), 1.0);
        for (int o = bI + 1; o < N; ++o) pair_block(pair_index(o, bI), 1.0);
      } else {
        pair_block(pair_index(bJ, bI), -1.0);
      }
#pragma unroll
      for (int u = 0; u < 3; ++u)
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const double val = hacc[u][v] * P.gstd;
          const int x = 3 * bI + u, y = 3 * bJ + v;
          if (hs_ld > 0) {
            double *Hg = hess + (size_t)geom * hs_stride;
            Hg[x * hs_ld + y] = val;
            if (bI != bJ) Hg[y * hs_ld + x] = val;
          } else {
            hess[((size_t)x * X + y) * n + geom] = val;
            if (bI != bJ) hess[((size_t)y * X + x) * n + geom] = val;
          }
        }
    }
    if (hess && hs_ld > X) {
      double *Hg = hess + (size_t)geom * hs_stride;
      const int np = hs_ld - X;
      for (int i = t; i < X * np; i += G2_THREADS) Hg[(i / np) * hs_ld + X + i % np] = 0.0;
    }
    __syncthreads();
  }
}

// returns 0 on launch, 1 if the configuration is outside this kernel's envelope (the caller falls back to k_gdml_eval)
static int launch_gdml_eval2(const PotDev &P, int n, const double *r, double *V, double *grad, double *hess, cudaStream_t st,
                             int hs_ld, size_t hs_stride) {
  const int N = P.n_atoms, D = P.n_desc;
  if (N > G2_MAXN || N < 2) return 1;
  if ((D * sizeof(double)) % 16 != 0) return 1;            // bulk copies need 16-byte rows
  const Gdml2Layout L = make_gdml2_layout(N, D);
  const size_t smem = sizeof(double) * (size_t)L.total;
  if (smem > 113 * 1024) return 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = sms * 2;
  if (grid > n) grid = n;
  if (N <= 17 && D <= 160) {
    if (cudaFuncSetAttribute(k_gdml_eval2<16, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    k_gdml_eval2<16, 5><<<grid, G2_THREADS, smem, st>>>(P, n, r, V, grad, hess, L, hs_ld, hs_stride);
  } else {
    if (cudaFuncSetAttribute(k_gdml_eval2<20, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
    k_gdml_eval2<20, 7><<<grid, G2_THREADS, smem, st>>>(P, n, r, V, grad, hess, L, hs_ld, hs_stride);
  }
  return 0;
}

}  // namespace sc
