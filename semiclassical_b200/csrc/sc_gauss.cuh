// sc_gauss.cuh -- wavefunction diagnostics of the Herman-Kluk propagator (propagators.py:657-782):
//   coefficients()   v_i = C_i e^{i S_i} <q_i,p_i|phi(0)> / ((2 pi)^d n P_i)                       :657-686
//   norm()           |psi|^2 = sum_ij conj(v_i) <q_i,p_i,Gt|q_j,p_j,Gt> v_j   (all pairs, O(n^2))   :734-782, 230-237
//   wavefunction(x)  psi(x_k) = sum_i v_i <x_k|q_i,p_i,Gt>                                          :688-732, 271-290
//
// Both sums have the form
//     out_i = sum_j coef_j exp( alpha_i + alphaJ_j + a_i . r_j  +  i (gamma_i + beta_j + a_i . s_j) )
// once the quadratic forms of the Gaussian exponents are expanded (bra-only terms, ket-only terms, one bilinear term
// each for the real and the imaginary part), so the pair loop is two real GEMMs  A R^T,  A S^T  (bras x K) (K x kets)
// with K = 2d (norm) or d (wavefunction) followed by exp / sincos per pair -- a dense contraction that runs on the
// FP64 tensor pipe (mma.sync.m8n8k4.f64), the transcendental epilogue on the accumulator fragments.
//   norm:          a_i = (q_i, p_i)   r_j = (A q_j, B p_j)   s_j = ((1 - C) p_j, -C^T q_j)
//                  alpha = alphaJ = -1/2 (q A q + p B p)   gamma_i = q_i C p_i   beta_j = -p_j.q_j + q_j C p_j
//                  with A = Gi (Gi+Gj)^-1 Gj, B = (Gi+Gj)^-1, C = Gj (Gi+Gj)^-1 for Gi = Gj = Gamma_t (:174-179)
//   wavefunction:  a_k = x_k   r_i = Gt q_i   s_i = p_i   alpha_k = -1/2 x_k Gt x_k   alphaJ_i = -1/2 q_i Gt q_i
//                  gamma_k = 0   beta_i = -p_i.q_i   coef_i = v_i (det Gt / pi^rank)^(1/4)
// One CTA (8 warps) owns 64 bras and walks over all kets in tiles of 32: every out_i has exactly one writer and a
// fixed summation order (deterministic, no atomics).
#pragma once
#include "sc_device.cuh"

namespace sc {

constexpr int GS_TI = 64, GS_TJ = 32, GS_THREADS = 256;
constexpr int SC_MAX_DIM_DEV = 96;     // = SC_MAX_DIM of the C ABI

__host__ __device__ constexpr int gs_ld(int kp) { return kp % 16 == 4 || kp % 16 == 12 ? kp : ((kp + 4) % 16 == 4 || (kp + 4) % 16 == 12 ? kp + 4 : kp + 8); }

__device__ __forceinline__ void gs_dmma(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// v_i = sign_i c_i exp(i S_i) wvi_i / ntraj_norm      (wvi = <q_i,p_i|phi(0)> / (P_i (2 pi)^d))
__global__ void k_coefficients(EngDev E, double inv_norm, double2 *__restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E.n) return;
  const double S = E.rec[(size_t)i * E.rs + 2 * E.d];
  const double2 c = E.c[i];
  const double sg = E.sign[i] * inv_norm;
  v[i] = cmul(cmul(make_double2(sg * c.x, sg * c.y), cexp(0.0, S)), E.wvi[i]);
}

// norm(): per-trajectory vectors and scalars.  A, B, C: (d x d) row-major in global memory.  One CTA per trajectory.
__global__ void __launch_bounds__(128)
k_gauss_prep_norm(EngDev E, const double *__restrict__ A, const double *__restrict__ B, const double *__restrict__ C, int kp,
                  double *__restrict__ a, double *__restrict__ r, double *__restrict__ s, double *__restrict__ alpha,
                  double *__restrict__ beta, double *__restrict__ gamma) {
  __shared__ double q[SC_MAX_DIM_DEV], p[SC_MAX_DIM_DEV], red[3][4];
  const int i = blockIdx.x, t = threadIdx.x, d = E.d, lane = t & 31, w = t >> 5;
  const double *rec = E.rec + (size_t)i * E.rs;
  if (t < d) { q[t] = rec[t]; p[t] = rec[d + t]; }
  __syncthreads();
  double v[3] = {0.0, 0.0, 0.0};
  if (t < d) {
    double aq = 0.0, bp = 0.0, cp = 0.0, ctq = 0.0;
    for (int j = 0; j < d; ++j) {
      aq += A[t * d + j] * q[j];
      bp += B[t * d + j] * p[j];
      cp += C[t * d + j] * p[j];
      ctq += C[j * d + t] * q[j];
    }
    double *ai = a + (size_t)i * kp, *ri = r + (size_t)i * kp, *si = s + (size_t)i * kp;
    ai[t] = q[t]; ai[d + t] = p[t];
    ri[t] = aq; ri[d + t] = bp;
    si[t] = p[t] - cp; si[d + t] = -ctq;
    v[0] = -0.5 * (q[t] * aq + p[t] * bp);
    v[1] = q[t] * cp;
    v[2] = -p[t] * q[t];
  }
  if (t >= 2 * d && t < kp) {                                  // zero padding of the K dimension
    a[(size_t)i * kp + t] = 0.0; r[(size_t)i * kp + t] = 0.0; s[(size_t)i * kp + t] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k][w] = v[k];
  }
  __syncthreads();
  if (t == 0) {
    const double s0 = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    const double s1 = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    const double s2 = (red[2][0] + red[2][1]) + (red[2][2] + red[2][3]);
    alpha[i] = s0;
    gamma[i] = s1;
    beta[i] = s2 + s1;
  }
}

// wavefunction(): kets (trajectories): r_i = G q_i, s_i = p_i, alphaJ_i, beta_i.  One CTA per trajectory.
__global__ void __launch_bounds__(128)
k_gauss_prep_wf_kets(EngDev E, const double *__restrict__ G, int kp, double *__restrict__ r, double *__restrict__ s,
                     double *__restrict__ alphaJ, double *__restrict__ beta) {
  __shared__ double q[SC_MAX_DIM_DEV], red[2][4];
  const int i = blockIdx.x, t = threadIdx.x, d = E.d, lane = t & 31, w = t >> 5;
  const double *rec = E.rec + (size_t)i * E.rs;
  if (t < d) q[t] = rec[t];
  __syncthreads();
  double v[2] = {0.0, 0.0};
  if (t < d) {
    double gq = 0.0;
    for (int j = 0; j < d; ++j) gq += G[t * d + j] * q[j];
    const double pt = rec[d + t];
    r[(size_t)i * kp + t] = gq;
    s[(size_t)i * kp + t] = pt;
    v[0] = -0.5 * q[t] * gq;
    v[1] = -pt * q[t];
  }
  if (t >= d && t < kp) { r[(size_t)i * kp + t] = 0.0; s[(size_t)i * kp + t] = 0.0; }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k][w] = v[k];
  }
  __syncthreads();
  if (t == 0) {
    alphaJ[i] = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    beta[i] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
  }
}

// wavefunction(): bras (grid points, x is (d, nx) like the reference's argument): a_k = x_k, alpha_k = -1/2 x G x
__global__ void __launch_bounds__(128)
k_gauss_prep_wf_bras(int d, int nx, const double *__restrict__ x, const double *__restrict__ G, int kp, double *__restrict__ a,
                     double *__restrict__ alpha, double *__restrict__ gamma) {
  __shared__ double xs[SC_MAX_DIM_DEV], red[4];
  const int k = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t < d) xs[t] = x[(size_t)t * nx + k];
  __syncthreads();
  double v = 0.0;
  if (t < d) {
    double gx = 0.0;
    for (int j = 0; j < d; ++j) gx += G[t * d + j] * xs[j];
    a[(size_t)k * kp + t] = xs[t];
    v = -0.5 * xs[t] * gx;
  }
  if (t >= d && t < kp) a[(size_t)k * kp + t] = 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (t == 0) {
    alpha[k] = (red[0] + red[1]) + (red[2] + red[3]);
    gamma[k] = 0.0;
  }
}

__device__ __forceinline__ void gs_cp_async16(void *smem_dst, const void *gsrc, bool valid) {
  // 16-byte asynchronous copy; !valid: zero fill (source size 0)
  const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(sz) : "memory");
}

// out_i = sum_j coef_j exp(alpha_i + alphaJ_j + a_i.r_j + i (gamma_i + beta_j + a_i.s_j));  kp = padded K (multiple of 4)
// The ket tiles (R, S rows and the ket scalars) are double buffered: tile j + 1 is fetched with cp.async while tile j
// is contracted and exponentiated (the synchronous version stalled 31 % of the time on these loads).
__global__ void __launch_bounds__(GS_THREADS, 1)
k_gauss_sum(int n_bra, int n_ket, int kp, const double *__restrict__ a, const double *__restrict__ alpha,
            const double *__restrict__ gamma, const double *__restrict__ r, const double *__restrict__ s,
            const double *__restrict__ alphaJ, const double *__restrict__ beta, const double2 *__restrict__ coef,
            double2 *__restrict__ out) {
  extern __shared__ __align__(16) double gsm[];
  const int ld = gs_ld(kp);
  double *As = gsm;                            // [64][ld]
  double *RS = As + GS_TI * ld;                // 2 buffers x { R [32][ld], S [32][ld] }
  double *KJ = RS + 4 * GS_TJ * ld;            // 2 buffers x [4][32]: alphaJ, beta, Re coef, Im coef
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int fr = lane >> 2, fc = lane & 3;
  const int i0 = blockIdx.x * GS_TI;
  for (int e = t; e < GS_TI * kp; e += GS_THREADS) {
    const int row = e / kp, k = e - row * kp;
    As[row * ld + k] = (i0 + row < n_bra) ? a[(size_t)(i0 + row) * kp + k] : 0.0;
  }
  const int irow = i0 + 8 * w + fr;                            // the bra this thread's accumulator rows belong to
  const double al_i = irow < n_bra ? alpha[irow] : 0.0, ga_i = irow < n_bra ? gamma[irow] : 0.0;
  double2 sum = make_double2(0.0, 0.0);
  const double *Af = As + (8 * w + fr) * ld + fc;
  const int nk = kp >> 2, kp2 = kp >> 1;                       // 16-byte pieces per row
  auto fetch = [&](int j0, int buf) {
    double *Rs = RS + buf * 2 * GS_TJ * ld, *Ss = Rs + GS_TJ * ld, *kj = KJ + buf * 128;
    for (int e = t; e < GS_TJ * kp2; e += GS_THREADS) {
      const int row = e / kp2, k = 2 * (e - row * kp2);
      const bool ok = j0 + row < n_ket;
      const size_t src = (size_t)(ok ? j0 + row : 0) * kp + k;
      gs_cp_async16(Rs + row * ld + k, r + src, ok);
      gs_cp_async16(Ss + row * ld + k, s + src, ok);
    }
    if (t < GS_TJ) {
      const bool ok = j0 + t < n_ket;
      const double2 cf = ok ? coef[j0 + t] : make_double2(0.0, 0.0);
      kj[t] = ok ? alphaJ[j0 + t] : 0.0;
      kj[32 + t] = ok ? beta[j0 + t] : 0.0;
      kj[64 + t] = cf.x;
      kj[96 + t] = cf.y;
    }
  };
  if (n_ket > 0) fetch(0, 0);
  int buf = 0;
  for (int j0 = 0; j0 < n_ket; j0 += GS_TJ, buf ^= 1) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                           // tile j0 has landed; everybody is done with tile j0 - 32
    if (j0 + GS_TJ < n_ket) fetch(j0 + GS_TJ, buf ^ 1);
    const double *Rs = RS + buf * 2 * GS_TJ * ld, *Ss = Rs + GS_TJ * ld, *kj = KJ + buf * 128;
    double accR[4][2], accS[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) accR[nt][0] = accR[nt][1] = accS[nt][0] = accS[nt][1] = 0.0;
    const double *Rf = Rs + fr * ld + fc, *Sf = Ss + fr * ld + fc;
#pragma unroll 2
    for (int kk = 0; kk < nk; ++kk) {
      const double af = Af[4 * kk];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        gs_dmma(accR[nt][0], accR[nt][1], af, Rf[nt * 8 * ld + 4 * kk]);
        gs_dmma(accS[nt][0], accS[nt][1], af, Sf[nt * 8 * ld + 4 * kk]);
      }
    }
    // epilogue on the accumulator fragments: element (row fr, column 8 nt + 2 fc + e)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * nt + 2 * fc + e;
        const double re = accR[nt][e] + al_i + kj[j], im = accS[nt][e] + ga_i + kj[32 + j];
        const double2 wv = cexp(re, im);
        const double cx = kj[64 + j], cy = kj[96 + j];
        sum.x += cx * wv.x - cy * wv.y;
        sum.y += cx * wv.y + cy * wv.x;
      }
    }
  }
  // the 4 lanes of a fragment row hold disjoint ket subsets of the same bra
  sum.x += __shfl_xor_sync(0xffffffffu, sum.x, 1);
  sum.y += __shfl_xor_sync(0xffffffffu, sum.y, 1);
  sum.x += __shfl_xor_sync(0xffffffffu, sum.x, 2);
  sum.y += __shfl_xor_sync(0xffffffffu, sum.y, 2);
  if (fc == 0 && irow < n_bra) out[irow] = sum;
}

// norm^2 = Re sum_i conj(v_i) out_i : one CTA, fixed order
__global__ void __launch_bounds__(256)
k_gauss_dot(int n, const double2 *__restrict__ v, const double2 *__restrict__ o, double *__restrict__ res) {
  __shared__ double red[2][8];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  double sx = 0.0, sy = 0.0;
  for (int i = t; i < n; i += 256) {
    const double2 a = v[i], b = o[i];
    sx += a.x * b.x + a.y * b.y;
    sy += a.x * b.y - a.y * b.x;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, off);
    sy += __shfl_xor_sync(0xffffffffu, sy, off);
  }
  if (lane == 0) { red[0][w] = sx; red[1][w] = sy; }
  __syncthreads();
  if (t == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 8; ++k) { a += red[0][k]; b += red[1][k]; }
    res[0] = a;
    res[1] = b;
  }
}

static inline size_t gs_smem_bytes(int kp) { return sizeof(double) * ((size_t)(GS_TI + 4 * GS_TJ) * gs_ld(kp) + 2 * 4 * 32); }

}  // namespace sc
