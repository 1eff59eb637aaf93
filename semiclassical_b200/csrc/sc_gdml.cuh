// sc_gdml.cuh -- sGDML energy / gradient / analytic Hessian (reference: gdml_predictor.py:140-250 behind
// MolecularGDMLPotential.harmonic_approximation, potentials.py:669-699).
//
// One CTA (8 warps) evaluates one geometry at a time.  The training matrices xs_train and Jx_alphas (M x D fp64,
// M = n_train * n_perms) are streamed through shared memory in tiles of 8 training points with 1-D bulk
// asynchronous copies (cp.async.bulk -> mbarrier, the TMA engine), double buffered; they are shared by every CTA
// and stay L2 resident.  Restructuring relative to the reference's dense einsums:
//   * the descriptor Jacobian J (D x 3N) has 6 non-zeros per row (the two atoms of the pair): XJ = x_diffs J and
//     AJ = A J are sums over the N-1 pairs of an atom, never dense D x 3N contractions
//   * the kernel sum over training points of the Hessian,
//         sum_m  w1_m XJ_mx XJ_my - ef_m (AJ_mx XJ_my + XJ_mx AJ_my),    w1 = ef XA q / |x - x_m|,
//     is G + G^T with G = a^T XJ, a_m = w1_m/2 XJ_m - ef_m AJ_m: one rank-M update of the upper-triangular
//     3x3 atom blocks held in registers (17 x 18 / 2 = 153 blocks for the coumarin-sized model)
//   * J^T J and the second-derivative scatter terms h1, h2 (gdml_predictor.py:205-244) touch only the 3x3
//     blocks (k,k), (l,l), (k,l), (l,k) of pair n = (k,l): one 3x3 block B_n per pair
// Summation order over the training points is m-ascending inside a warp slot and slot-ascending at the end;
// the reference's own result moves by ~3e-9 (absolute) under a permutation of the training set.
#pragma once
#include <cstdint>
#include <cstdlib>

#include "sc_device.cuh"

namespace sc {

constexpr int GDML_TM = 8;        // training points per tile == warps per CTA
constexpr int GDML_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct GdmlLayout {
  int xt, at, xd, xj, aj, xs, jv, gx, gxw, sc, rp, total;  // offsets in doubles
  int Xp;
};

__host__ __device__ inline GdmlLayout make_gdml_layout(int N, int D) {
  GdmlLayout L;
  const int X = 3 * N;
  L.Xp = (X + 1) & ~1;
  int o = 0;
  L.xt = o; o += 2 * GDML_TM * D;     // training descriptors, double buffered
  L.at = o; o += 2 * GDML_TM * D;     // Jx_alphas rows, double buffered
  L.xd = o; o += GDML_TM * D;         // x - x_m of the current tile
  L.xj = o; o += GDML_TM * L.Xp;
  L.aj = o; o += GDML_TM * L.Xp;
  L.xs = o; o += (D + 1) & ~1;
  L.jv = o; o += 3 * D + (D & 1);
  L.gx = o; o += (D + 1) & ~1;
  L.gxw = o; o += GDML_TM * D;        // per-warp partial dE/dx_desc
  L.sc = o; o += 6 * GDML_TM;         // per-slot scalars: ef, w1/2, energy partial, sum ef XA partial
  L.rp = o; o += L.Xp;
  L.total = (o + 1) & ~1;
  return L;
}

// pair index of atoms (i, j), i > j, in the lower-triangular enumeration of torch.tril_indices(N, N, -1)
__device__ __forceinline__ int pair_index(int i, int j) { return i * (i - 1) / 2 + j; }

__global__ void __launch_bounds__(GDML_THREADS)
k_gdml_eval(PotDev P, int n, const double *__restrict__ r, double *__restrict__ V, double *__restrict__ grad,
            double *__restrict__ hess, GdmlLayout L, int use_bulk, int hs_ld, size_t hs_stride) {
  // hs_ld > 0: the Hessian of geometry g goes to hess + g hs_stride as a row-major image with leading dimension hs_ld
  // (zero padding columns): the stream image k_rk4_stream consumes (sc_stream.cuh); else batch-last (X, X, n)
  extern __shared__ __align__(16) double gsm[];
  __shared__ __align__(8) uint64_t bars[2];
  const int N = P.n_atoms, M = P.n_train, D = P.n_desc, X = 3 * N, Xp = L.Xp;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const double q = sqrt(5.0) / P.sig, q2 = q * q, q4_3 = q2 * q2 / 3.0;
  double *xtb = gsm + L.xt, *atb = gsm + L.at, *xd = gsm + L.xd, *XJ = gsm + L.xj, *AJ = gsm + L.aj;
  double *xs = gsm + L.xs, *jv = gsm + L.jv, *gx = gsm + L.gx, *gxw = gsm + L.gxw, *sc = gsm + L.sc, *rp = gsm + L.rp;
  const int ntiles = (M + GDML_TM - 1) / GDML_TM;
  // upper-triangular atom block (I <= J) owned by this thread
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
  __syncthreads();
  uint32_t phase[2] = {0u, 0u};

  auto issue_tile = [&](int tile) {
    const int b = tile & 1, m0 = tile * GDML_TM;
    const int rows = (M - m0 < GDML_TM) ? M - m0 : GDML_TM;
    if (use_bulk) {
      if (t == 0) {
        const uint32_t bytes = (uint32_t)(rows * D * sizeof(double));
        mbar_expect_tx(&bars[b], 2 * bytes);
        bulk_g2s(xtb + b * GDML_TM * D, P.xs_train + (size_t)m0 * D, bytes, &bars[b]);
        bulk_g2s(atb + b * GDML_TM * D, P.jx_alphas + (size_t)m0 * D, bytes, &bars[b]);
      }
    } else {
      for (int i = t; i < rows * D; i += GDML_THREADS) {
        xtb[b * GDML_TM * D + i] = P.xs_train[(size_t)m0 * D + i];
        atb[b * GDML_TM * D + i] = P.jx_alphas[(size_t)m0 * D + i];
      }
    }
  };

  for (int geom = blockIdx.x; geom < n; geom += gridDim.x) {
    // ---- descriptor and the non-zeros of its Jacobian
    for (int x = t; x < X; x += GDML_THREADS) rp[x] = r[(size_t)x * n + geom];
    __syncthreads();
    issue_tile(0);
    for (int p = t; p < D; p += GDML_THREADS) {
      // pair p = (i, j), i > j
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
    for (int i = t; i < GDML_TM * D; i += GDML_THREADS) gxw[i] = 0.0;
    if (t < 2 * GDML_TM) sc[2 * GDML_TM + t] = 0.0;   // energy and sum ef XA partials
    double hacc[3][3];
#pragma unroll
    for (int u = 0; u < 3; ++u)
#pragma unroll
      for (int v = 0; v < 3; ++v) hacc[u][v] = 0.0;
    __syncthreads();

    // ---- kernel sum over the training points
    for (int tile = 0; tile < ntiles; ++tile) {
      const int b = tile & 1, m0 = tile * GDML_TM;
      if (tile + 1 < ntiles) issue_tile(tile + 1);   // the other buffer was released by the barrier below
      if (use_bulk) {
        mbar_wait(&bars[b], phase[b]);
        phase[b] ^= 1u;
      } else {
        __syncthreads();
      }
      const int m = m0 + warp;
      const bool live = m < M;
      const double *xt = xtb + (b * GDML_TM + warp) * D, *A = atb + (b * GDML_TM + warp) * D;
      double *xdw = xd + warp * D;
      double ef = 0.0, hw1 = 0.0;
      if (live) {
        double n2 = 0.0, xa = 0.0;
        for (int p = lane; p < D; p += 32) {
          const double df = xs[p] - xt[p];
          xdw[p] = df;
          n2 = fma(df, df, n2);
          xa = fma(df, A[p], xa);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          n2 += __shfl_xor_sync(0xffffffffu, n2, o);
          xa += __shfl_xor_sync(0xffffffffu, xa, o);
        }
        const double xn = sqrt(n2);
        ef = q4_3 * exp(-q * xn);
        const double k1 = ef * (1.0 + q * xn) / q2;
        hw1 = 0.5 * ef * xa * q / xn;
        const double efxa = ef * xa;
        double *gw = gxw + warp * D;
        for (int p = lane; p < D; p += 32) gw[p] += k1 * A[p] - efxa * xdw[p];
        if (lane == 0) {
          sc[2 * GDML_TM + warp] += k1 * xa;
          sc[3 * GDML_TM + warp] += efxa;
        }
        __syncwarp();
        // XJ_m, AJ_m: sums over the N-1 pairs of the atom that owns coordinate x
        for (int x = lane; x < X; x += 32) {
          const int I = x / 3, u = x - 3 * I;
          double sx = 0.0, sa = 0.0;
          for (int o = 0; o < I; ++o) {          // pairs (I, o): this atom is the first index, J = -jv
            const int p = pair_index(I, o);
            const double jj = jv[3 * p + u];
            sx = fma(-xdw[p], jj, sx);
            sa = fma(-A[p], jj, sa);
          }
          for (int o = I + 1; o < N; ++o) {      // pairs (o, I): second index, J = +jv
            const int p = pair_index(o, I);
            const double jj = jv[3 * p + u];
            sx = fma(xdw[p], jj, sx);
            sa = fma(A[p], jj, sa);
          }
          XJ[warp * Xp + x] = sx;
          AJ[warp * Xp + x] = hw1 * sx - ef * sa;     // a_m
        }
      } else {
        for (int x = lane; x < X; x += 32) { XJ[warp * Xp + x] = 0.0; AJ[warp * Xp + x] = 0.0; }
      }
      __syncthreads();
      if (bI >= 0) {
#pragma unroll 2
        for (int mm = 0; mm < GDML_TM; ++mm) {
          const double *xj = XJ + mm * Xp, *aj = AJ + mm * Xp;
          double xi[3], ai[3], xjv[3], ajv[3];
#pragma unroll
          for (int u = 0; u < 3; ++u) { xi[u] = xj[3 * bI + u]; ai[u] = aj[3 * bI + u]; xjv[u] = xj[3 * bJ + u]; ajv[u] = aj[3 * bJ + u]; }
#pragma unroll
          for (int u = 0; u < 3; ++u)
#pragma unroll
            for (int v = 0; v < 3; ++v) hacc[u][v] = fma(ai[u], xjv[v], fma(xi[u], ajv[v], hacc[u][v]));
        }
      }
      __syncthreads();
    }

    // ---- reductions over the warp slots (fixed order)
    for (int p = t; p < D; p += GDML_THREADS) {
      double s = 0.0;
      for (int w = 0; w < GDML_TM; ++w) s += gxw[w * D + p];
      gx[p] = s;
    }
    double en = 0.0, sumefxa = 0.0;
    for (int w = 0; w < GDML_TM; ++w) { en += sc[2 * GDML_TM + w]; sumefxa += sc[3 * GDML_TM + w]; }
    __syncthreads();
    if (t == 0) V[geom] = en * P.gstd + P.e0 - P.origin;
    if (grad) {
      for (int x = t; x < X; x += GDML_THREADS) {
        const int I = x / 3, u = x - 3 * I;
        double s = 0.0;
        for (int o = 0; o < I; ++o) { const int p = pair_index(I, o); s = fma(-gx[p], jv[3 * p + u], s); }
        for (int o = I + 1; o < N; ++o) { const int p = pair_index(o, I); s = fma(gx[p], jv[3 * p + u], s); }
        grad[(size_t)x * n + geom] = s * P.gstd;
      }
    }
    if (hess && bI >= 0) {
      // pair blocks B_p[u][v] = -sum(ef XA) jv_u jv_v + 3 gx jv_u jv_v / x + delta_uv (-gx x^3):
      // added to the diagonal atom blocks of both atoms, subtracted from the off-diagonal ones
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
          // diagonal blocks hold G + G^T already symmetrised element-wise: (u,v) and (v,u) are both computed
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
      for (int i = t; i < X * np; i += GDML_THREADS) Hg[(i / np) * hs_ld + X + i % np] = 0.0;
    }
    __syncthreads();
  }
}

// returns 0 on launch, 1 if the configuration is outside the kernel's envelope
static int launch_gdml_eval2(const PotDev &P, int n, const double *r, double *V, double *grad, double *hess, cudaStream_t st,
                             int hs_ld, size_t hs_stride);   // sc_gdml2.cuh

static int launch_gdml_eval(const PotDev &P, int n, const double *r, double *V, double *grad, double *hess, cudaStream_t st,
                            int hs_ld = 0, size_t hs_stride = 0) {
  const int N = P.n_atoms, D = P.n_desc;
  // second-generation kernel (register-resident Jacobian, DMMA rank-M update) inside its envelope, unless SC_GDML_V1=1
  if (!getenv("SC_GDML_V1") && launch_gdml_eval2(P, n, r, V, grad, hess, st, hs_ld, hs_stride) == 0) return 0;
  if (N * (N + 1) / 2 > GDML_THREADS) return 1;
  const GdmlLayout L = make_gdml_layout(N, D);
  const size_t smem = sizeof(double) * (size_t)L.total;
  if (smem > 227 * 1024) return 1;
  if (cudaFuncSetAttribute(k_gdml_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  int grid = sms * per_sm;
  if (grid > n) grid = n;
  const int use_bulk = ((D * sizeof(double)) % 16 == 0) ? 1 : 0;
  k_gdml_eval<<<grid, GDML_THREADS, smem, st>>>(P, n, r, V, grad, hess, L, use_bulk, hs_ld, hs_stride);
  return 0;
}

}  // namespace sc
