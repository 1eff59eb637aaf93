// sc_mma.cuh -- FP64 tensor-core variant of the fused Herman-Kluk step for large d (17 <= d <= 62).
//
// One CTA per trajectory, resident for K time steps.  Per RK4 stage the monodromy products
//     [dMpq | dMpp] = -H(q_s) [Mqq | Mqp]_s            (propagators.py:347, 357; 4 d^3 flops per stage)
// run on the FP64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), A = H (row-major, shared), B = U_s
// (shared), accumulators in registers.  tcgen05/TMEM has no f64 kind, so warp-level DMMA is the tensor path
// for this problem.  Each thread owns the C-fragment elements of its warp tile for the whole step, so the RK4
// accumulators (R1, R2 of sc_kernels.cuh) never leave registers.
//
// Shared-memory plan for d = 60 (bytes): Ub, Vb, Us 3 x 57 600 + H 64 x 60 x 8 = 30 720 + vectors ~= 210 KB.
//   Ub, Vb : ld = 2d (+pad so that ld mod 16 == 8): 128-bit owner accesses are conflict-free
//   Us     : same ld, columns XOR-swizzled by ((row>>1)&1)<<2 so that the B-fragment loads (4 k-rows x 8 columns
//            per quarter) are conflict-free as well
//   H      : ld mod 16 in {4, 12}: conflict-free A-fragment loads
#pragma once
#include "sc_kernels.cuh"

namespace sc {

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int swz(int row) { return ((row >> 1) & 1) << 2; }

// separable potentials: thread t < d evaluates its own mode, no cross-thread dependence
__device__ __forceinline__ double pot_local(const PotDev &P, int t, double r, double *g, double *H, int ldh) {
  double v, hd;
  if (P.type == POT_MORSE) {
    if (P.all_harmonic) {
      const double w2 = P.omega[t] * P.omega[t];
      v = 0.5 * w2 * r * r;
      g[t] = w2 * r;
      hd = w2;
    } else {
      const double a = P.a[t], D = P.D[t];
      const double e = exp(-a * r);
      v = D * (1.0 - e) * (1.0 - e);
      g[t] = 2.0 * a * D * e * (1.0 - e);
      hd = 2.0 * a * a * D * e * (2.0 * e - 1.0);
    }
  } else {
    const double eps = P.eps[t], b = P.b[t];
    const double e1 = exp(-b * r), e2 = exp(-2.0 * b * r);
    v = eps / (2.0 * b * b) * (1.0 - e1) * (1.0 - e1) + (1.0 - eps) * 0.5 * r * r;
    g[t] = eps / b * (e1 - e2) + (1.0 - eps) * r;
    hd = eps * (2.0 * e2 - e1) + (1.0 - eps);
  }
  H[t * ldh + t] = hd;
  if (t == 0) v -= P.origin;
  return v;
}

// optional phase timing (compile with -DSC_PHASE_TIMING): thread 0 of CTA 0 accumulates clock64() deltas
#ifdef SC_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define PT_DECL long long pt_last = clock64();
#define PT(idx) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long now_ = clock64(); g_phase_cycles[idx] += (unsigned long long)(now_ - pt_last); pt_last = now_; } } while (0)
#else
#define PT_DECL
#define PT(idx) do { } while (0)
#endif

constexpr int MMA_KMAX = 128;   // time steps per launch (size of the shared correlation accumulators)

// separable potentials without side effects: value, gradient and Hessian diagonal of mode t at r
__device__ __forceinline__ double pot_local_vals(const PotDev &P, int t, double r, double &gout, double &hd) {
  double v;
  if (P.type == POT_MORSE) {
    if (P.all_harmonic) {
      const double w2 = P.omega[t] * P.omega[t];
      v = 0.5 * w2 * r * r;
      gout = w2 * r;
      hd = w2;
    } else {
      const double a = P.a[t], D = P.D[t];
      const double e = exp(-a * r);
      v = D * (1.0 - e) * (1.0 - e);
      gout = 2.0 * a * D * e * (1.0 - e);
      hd = 2.0 * a * a * D * e * (2.0 * e - 1.0);
    }
  } else {
    const double eps = P.eps[t], b = P.b[t];
    const double e1 = exp(-b * r), e2 = exp(-2.0 * b * r);
    v = eps / (2.0 * b * b) * (1.0 - e1) * (1.0 - e1) + (1.0 - eps) * 0.5 * r * r;
    gout = eps / b * (e1 - e2) + (1.0 - eps) * r;
    hd = eps * (2.0 * e2 - e1) + (1.0 - eps);
  }
  if (t == 0) v -= P.origin;
  return v;
}

// H = Q diag(h) Q^T of the rotated-AS potential (potentials of the dense parity fixtures) on the tensor pipe:
// work item = (row tile, group of 4 column tiles); A fragment = Q[i][k] h[k], B fragment = Q[j][k], both straight from
// the L1/L2-resident Q (28.8 KB at d = 60).  The scalar version (pot_eval) needed d^3 global loads per stage.
// hk: shared vector of the d inner second derivatives.  Ends with a CTA barrier.
template <int NW>
__device__ __forceinline__ void rotated_hessian_mma(const PotDev &P, const double *hk, double *H, int ldh, int warp, int lane) {
  const int d = P.d;
  const int fr = lane >> 2, fc = lane & 3;
  const int mt_n = (d + 7) >> 3, nk = (d + 3) >> 2, ngr = (mt_n + 3) >> 2;
  for (int item = warp; item < mt_n * ngr; item += NW) {
    const int mt = item / ngr, n0 = 4 * (item - mt * ngr);
    const double *ap = P.Q + (size_t)min(8 * mt + fr, d - 1) * d + fc;
    double c[4][2];
#pragma unroll
    for (int n = 0; n < 4; ++n) c[n][0] = c[n][1] = 0.0;
#pragma unroll 3
    for (int kk = 0; kk < nk; ++kk) {
      const bool kok = 4 * kk + fc < d;
      const double a = kok ? __ldg(ap + 4 * kk) * hk[4 * kk + fc] : 0.0;
#pragma unroll
      for (int n = 0; n < 4; ++n)
        if (n0 + n < mt_n) {
          const double b = kok ? __ldg(P.Q + (size_t)min(8 * (n0 + n) + fr, d - 1) * d + 4 * kk + fc) : 0.0;
          dmma884(c[n][0], c[n][1], a, b);
        }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int row = 8 * mt + fr, col = 8 * (n0 + n) + 2 * fc;
      if (row < d && n0 + n < mt_n) {
        if (col < d) H[row * ldh + col] = c[n][0];
        if (col + 1 < d) H[row * ldh + col + 1] = c[n][1];
      }
    }
  }
  __syncthreads();
}

// Dense-Gamma prefactor matrix on the tensor pipe (the DFMA version, prefactor_assemble in sc_device.cuh, took 80 % of
// the step of a dense 60-mode model):
//     Cm = 1/2 [ L1 Mqq R1 + L2 Mpp R2 - i L1 Mqp R2 + i L2 Mpq R1 ]            (propagators.py:969-994)
// as four pairs of products  T = M_blk R_blk (d x dr, to shared memory),  C += +-L_blk T (dr x dr, accumulators
// in registers across the four blocks).  8 x 8 output tiles are dealt round-robin to the NW warps; M_blk comes from the
// monodromy slabs in shared memory; the constant factor matrices L_blk, R_blk are staged per block into S (the stage-
// operand region, dead between stage 4 and the next step) with coalesced loads -- fetching their fragments straight
// from L2 made the assembly latency bound (1.7 MB of fragment loads per trajectory-step instead of 0.2 MB).
// T: >= ((d + 7) & ~7) rows x ldt; ldt and the staging leading dimensions are = 4 or 12 mod 16 (conflict-free fragments).
__host__ __device__ constexpr int mma_frag_ld(int n) { return n % 16 == 4 || n % 16 == 12 ? n : (n % 4 == 0 ? ((n + 4) % 16 == 4 || (n + 4) % 16 == 12 ? n + 4 : n + 8) : mma_frag_ld((n + 3) & ~3)); }

template <int NW>
__device__ __forceinline__ void prefactor_assemble_mma(const EngDev &E, const double *Ub, const double *Vb, int ldu, double2 *Cm,
                                                       int ldc, double *T, int ldt, double *S, int t, int warp, int lane) {
  // work item = (row tile, group of 4 column tiles): one A fragment feeds 4 DMMAs.  At most 8 x 2 items per product.
  constexpr int MAXI = (16 + NW - 1) / NW;
  constexpr int TPT = 32 * NW;
  const int d = E.d, dr = E.dr;
  const int fr = lane >> 2, fc = lane & 3;
  const int mtA = (d + 7) >> 3, ntA = (dr + 7) >> 3, nk = (d + 3) >> 2, dk = 4 * nk;
  const int ngr = (ntA + 3) >> 2, nitemsA = mtA * ngr, nitemsC = ntA * ngr;
  const int ldr = mma_frag_ld(dr), ldl = mma_frag_ld(d);
  double *Rs = S, *Ls = S + dk * ldr;                            // R_blk (dk x dr, zero rows beyond d), L_blk (dr x dk)
  const bool kpad = (d & 3) != 0;                                // k-steps reach beyond d: the slab columns there are not zero
  double cre[MAXI][4][2], cim[MAXI][4][2];
#pragma unroll
  for (int i = 0; i < MAXI; ++i)
#pragma unroll
    for (int n = 0; n < 4; ++n) cre[i][n][0] = cre[i][n][1] = cim[i][n][0] = cim[i][n][1] = 0.0;
#pragma unroll
  for (int blk = 0; blk < 4; ++blk) {
    // blk 0: Mqq (L1,R1) re ; 1: Mpp (L2,R2) re ; 2: Mqp (L1,R2) -im ; 3: Mpq (L2,R1) +im
    const double *Mb = ((blk == 0 || blk == 2) ? Ub : Vb) + ((blk == 1 || blk == 2) ? d : 0);
    const double *Lm = (blk == 0 || blk == 2) ? E.L1 : E.L2;
    const double *Rm = (blk == 0 || blk == 3) ? E.R1 : E.R2;
    const double sgn = (blk == 2) ? -1.0 : 1.0;
    for (int idx = t; idx < dk * dr; idx += TPT) {
      const int k = idx / dr, n = idx - k * dr;
      Rs[k * ldr + n] = k < d ? __ldg(Rm + idx) : 0.0;
    }
    for (int idx = t; idx < dr * dk; idx += TPT) {
      const int ap = idx / dk, a = idx - ap * dk;
      Ls[ap * ldl + a] = a < d ? sgn * __ldg(Lm + ap * d + a) : 0.0;
    }
    __syncthreads();
    // ---- T = M_blk R_blk.  Rows >= d of the last row tile repeat row d - 1 and columns >= dr of the last column tile
    // read the neighbouring row of Rs (finite; neither is stored)
    for (int item = warp; item < nitemsA; item += NW) {
      const int mt = item / ngr, n0 = 4 * (item - mt * ngr);
      const int row = min(8 * mt + fr, d - 1);
      const double *ap = Mb + row * ldu + fc;
      const double *bp = Rs + fc * ldr + 8 * n0 + fr;
      double c[4][2];
#pragma unroll
      for (int n = 0; n < 4; ++n) c[n][0] = c[n][1] = 0.0;
#pragma unroll 3
      for (int kk = 0; kk < nk; ++kk) {
        double a = ap[4 * kk];
        if (kpad && 4 * kk + fc >= d) a = 0.0;
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (n0 + n < ntA) dmma884(c[n][0], c[n][1], a, bp[4 * kk * ldr + 8 * n]);
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int col = 8 * (n0 + n) + 2 * fc;                        // columns >= dr would run into the next row
        double *tp = T + (8 * mt + fr) * ldt + col;
        if (8 * mt + fr < d) {                                        // the padding rows of T (= of H) stay zero
          if (col < dr) tp[0] = c[n][0];
          if (col + 1 < dr) tp[1] = c[n][1];
        }
      }
    }
    __syncthreads();
    // ---- C += +-L_blk T
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int item = warp + NW * i;
      if (item < nitemsC) {                                      // warp-uniform
        const int mt = item / ngr, n0 = 4 * (item - mt * ngr);
        const int row = min(8 * mt + fr, dr - 1);
        const double *ap = Ls + row * ldl + fc;
        const double *bp = T + fc * ldt + 8 * n0 + fr;
#pragma unroll 3
        for (int kk = 0; kk < nk; ++kk) {
          const double a = ap[4 * kk];
#pragma unroll
          for (int n = 0; n < 4; ++n)
            if (n0 + n < ntA) {
              if (blk < 2) dmma884(cre[i][n][0], cre[i][n][1], a, bp[4 * kk * ldt + 8 * n]);
              else dmma884(cim[i][n][0], cim[i][n][1], a, bp[4 * kk * ldt + 8 * n]);
            }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int item = warp + NW * i;
    if (item < nitemsC) {
      const int mt = item / ngr, n0 = 4 * (item - mt * ngr);
      const int row = 8 * mt + fr;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int col = 8 * (n0 + n) + 2 * fc;
        if (row < dr && n0 + n < ntA) {
          if (col < dr) Cm[row * ldc + col] = make_double2(0.5 * cre[i][n][0], 0.5 * cim[i][n][0]);
          if (col + 1 < dr) Cm[row * ldc + col + 1] = make_double2(0.5 * cre[i][n][1], 0.5 * cim[i][n][1]);
        }
      }
    }
  }
  __syncthreads();
}

template <int WM, int WN, int NWM, int NWN, int MC>
__global__ void __launch_bounds__(32 * NWM * NWN, 1)
k_hk_mma(EngDev E, PotDev P, double h, int nsteps, int step0, int nsteps_total, double *partials, SmemLayout L, int traj0,
         int ntw, double2 *__restrict__ cm, double *__restrict__ aux) {
  // cm != nullptr ("split" mode): the prefactor matrix of every (step, trajectory) of the window [traj0, traj0 + ntw)
  // goes to cm[(step ntw + tl) dr^2 ...] (LU column a, LU row b at [a dr + b]) and the overlap / action / energy sums
  // to aux[(step ntw + tl) 8 ...]; the determinants are taken by the batched DMMA LU (sc_lu_mma.cuh) and the branch
  // tracking + contributions by k_hk_finish -- the in-kernel register LU was 59 % of the fused step.
  const bool split = cm != nullptr;
  constexpr int NW = NWM * NWN, TPT = 32 * NW;
  extern __shared__ __align__(16) double smem[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int gg = blockIdx.x, NG = gridDim.x, gid = 0;
  double *Ub = smem + L.off_Ub, *Vb = smem + L.off_Vb, *Us = smem + L.off_Us, *H = smem + L.off_H;
  double *vec = smem + L.off_vec, *red = smem + L.off_red, *cacc = smem + L.off_acc;
  LuShared *lush = reinterpret_cast<LuShared *>(smem + L.off_lu);
  const int d = E.d, dr = E.dr, ldu = L.ldu, ldh = L.ldh, dp = L.dpad, W = 2 * d, NE = 2 * d * d;
  const int DK = (d + 3) & ~3;
  double *q = vec, *p = vec + dp, *qs = vec + 2 * dp, *g = vec + 3 * dp, *scr = vec + 4 * dp, *scr2 = vec + 5 * dp;
  double *dqv = vec + 6 * dp, *dpv = vec + 7 * dp, *hdv = vec + 8 * dp, *sacc = vec + 12 * dp, *se4 = vec + 13 * dp;
  double2 *Cm = reinterpret_cast<double2 *>(Us);
  const double im_t = (t < d) ? P.imass[t] : 0.0;
  const bool separable = (P.type == POT_MORSE || P.type == POT_NONHARMONIC);
  // warp tile origin and this thread's fragment coordinates
  const int m0 = (warp % NWM) * WM * 8, n0 = (warp / NWM) * WN * 8;
  const int fr = lane >> 2, fc = lane & 3;
  // swizzled B-fragment columns of this thread (loop invariant: the k-row parity bit is fc's); tiles that lie
  // entirely in the column padding read column 0 and produce ignored results
  int bcol[WN];
#pragma unroll
  for (int j = 0; j < WN; ++j) bcol[j] = ((n0 + 8 * j < W) ? (n0 + 8 * j + fr) : 0) ^ swz(fc);

  for (int i = t; i < 5 * nsteps; i += TPT) cacc[i] = 0.0;
  for (int i = t; i < ((d + 7) & ~7) * ldh; i += TPT) H[i] = 0.0;
  PT_DECL
  for (int tl = gg; tl < ntw; tl += NG) {
    const int traj = traj0 + tl;
    double *rec = E.rec + (size_t)traj * E.rs;
    if (t < d) { q[t] = rec[t]; p[t] = rec[d + t]; }
    double S = rec[2 * d];
    // zero the K-padding of the operands once
    for (int idx = t; idx < (DK - d) * ldu; idx += TPT) Us[d * ldu + idx] = 0.0;
    for (int idx = t; idx < NE; idx += TPT) {
      const int a = idx / W, b = idx % W;
      const double u = rec[E.qps + idx];
      Ub[a * ldu + b] = u;
      Us[a * ldu + (b ^ swz(a))] = u;
      Vb[a * ldu + b] = rec[E.qps + NE + idx];
    }
    double2 c2 = E.c2[traj], cc = E.c[traj];
    double sign = E.sign[traj];
    const double2 wvi = E.wvi[traj];
    __syncthreads();
    PT(0);

    for (int step = 0; step < nsteps; ++step) {
      double e4 = 0.0, accS = 0.0;
      double R1[WM][WN][2], R2[WM][WN][2];
      double qa = 0, pa = 0, qsa = 0, psa = 0, accq = 0, accp = 0;
      double vpart = 0.0;
      if (separable) {
        // (q, p) do not depend on the monodromy blocks: run their whole RK4 step now and keep the four
        // Hessian diagonals for the matrix stages
        if (t < d) {
          qa = q[t]; pa = p[t]; qsa = qa; psa = pa;
#pragma unroll
          for (int s = 1; s <= 4; ++s) {
            const double cnext = (s == 3) ? h : 0.5 * h;
            const double wgt = (s == 1 || s == 4) ? 1.0 : 2.0;
            double gt, hd;
            vpart = pot_local_vals(P, t, qsa, gt, hd);
            hdv[(s - 1) * dp + t] = hd;
            const double kq = psa * im_t, kp = -gt;
            const double tk = 0.5 * psa * psa * im_t;
            accS += wgt * (tk - vpart);
            if (s == 4) e4 = tk + vpart;
            accq += wgt * kq;
            accp += wgt * kp;
            if (s < 4) {
              qsa = qa + cnext * kq;
              psa = pa + cnext * kp;
            }
          }
          q[t] = qa + h / 6.0 * accq;
          p[t] = pa + h / 6.0 * accp;
          sacc[t] = accS;
          se4[t] = e4;
          H[t * ldh + t] = hdv[t];
        }
        __syncthreads();
      } else {
        if (t < d) { qa = q[t]; pa = p[t]; qsa = qa; psa = pa; qs[t] = qa; }
        // the dense prefactor assembly uses H as scratch: refill every step
        for (int i = t; i < ((d + 7) & ~7) * ldh; i += TPT) H[i] = 0.0;
        __syncthreads();
        vpart = pot_eval<TPT>(P, qs, g, H, ldh, scr, scr2, t, gid, true, true);
        if (P.type == POT_ROTATED_MORSE) rotated_hessian_mma<NW>(P, scr2, H, ldh, warp, lane);
      }
      PT(1);
#pragma unroll 1
      for (int s = 1; s <= 4; ++s) {
        const double cnext = (s == 3) ? h : 0.5 * h;
        const double wgt = (s == 1 || s == 4) ? 1.0 : 2.0;
        // ---- phase A: acc = H U_s on the tensor pipe (software-pipelined fragment loads)
        double acc[WM][WN][2];
#pragma unroll
        for (int i = 0; i < WM; ++i)
#pragma unroll
          for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        {
          const double *Ha = H + (m0 + fr) * ldh + fc;
          const double *Bp = Us + fc * ldu;
          const int ldu4 = 4 * ldu, ldh8 = 8 * ldh;
          double a[WM], b[WN];
#pragma unroll
          for (int i = 0; i < WM; ++i) a[i] = Ha[i * ldh8];
#pragma unroll
          for (int j = 0; j < WN; ++j) b[j] = Bp[bcol[j]];
          const int nk = DK >> 2;
#pragma unroll 5
          for (int k = 1; k <= nk; ++k) {
            double an[WM], bn[WN];
            if (k < nk) {
              Ha += 4;
              Bp += ldu4;
#pragma unroll
              for (int i = 0; i < WM; ++i) an[i] = Ha[i * ldh8];
#pragma unroll
              for (int j = 0; j < WN; ++j) bn[j] = Bp[bcol[j]];
            }
#pragma unroll
            for (int i = 0; i < WM; ++i)
#pragma unroll
              for (int j = 0; j < WN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
#pragma unroll
            for (int i = 0; i < WM; ++i) a[i] = an[i];
#pragma unroll
            for (int j = 0; j < WN; ++j) b[j] = bn[j];
          }
        }
        double kq = 0, kp = 0;
        if (!separable && t < d) {
          kq = psa * im_t;
          kp = -g[t];
          const double tk = 0.5 * psa * psa * im_t;
          accS += wgt * (tk - vpart);
          if (s == 4) e4 = tk + vpart;
        }
        __syncthreads();
        PT(2);
        // ---- phase B: RK4 accumulators (kv = -acc), operand of the next stage, vector part
#pragma unroll
        for (int i = 0; i < WM; ++i) {
          const int a = m0 + 8 * i + fr;
          const double ima = (a < d) ? P.imass[a] : 0.0;
          const int sw = swz(a);
#pragma unroll
          for (int j = 0; j < WN; ++j) {
            const int b = n0 + 8 * j + 2 * fc;
            if (a < d && b < W) {
              const double2 ub = *reinterpret_cast<const double2 *>(Ub + a * ldu + b);
              const double2 vb = *reinterpret_cast<const double2 *>(Vb + a * ldu + b);
              const double k0v = -acc[i][j][0], k1v = -acc[i][j][1];
              double2 un;
              if (s == 1) {
                R1[i][j][0] = k0v; R1[i][j][1] = k1v;
                R2[i][j][0] = 0.0; R2[i][j][1] = 0.0;
                un.x = ub.x + 0.5 * h * vb.x * ima;
                un.y = ub.y + 0.5 * h * vb.y * ima;
              } else if (s == 2) {
                un.x = ub.x + 0.5 * h * (vb.x + 0.5 * h * R1[i][j][0]) * ima;
                un.y = ub.y + 0.5 * h * (vb.y + 0.5 * h * R1[i][j][1]) * ima;
                R1[i][j][0] += k0v; R1[i][j][1] += k1v;
                R2[i][j][0] = k0v; R2[i][j][1] = k1v;
              } else if (s == 3) {
                un.x = ub.x + h * (vb.x + 0.5 * h * R2[i][j][0]) * ima;
                un.y = ub.y + h * (vb.y + 0.5 * h * R2[i][j][1]) * ima;
                R1[i][j][0] += k0v; R1[i][j][1] += k1v;
                R2[i][j][0] += k0v; R2[i][j][1] += k1v;
              } else {
                R2[i][j][0] += k0v; R2[i][j][1] += k1v;
                un.x = ub.x + h * vb.x * ima + (h * h / 6.0) * R1[i][j][0] * ima;
                un.y = ub.y + h * vb.y * ima + (h * h / 6.0) * R1[i][j][1] * ima;
                double2 vn;
                vn.x = vb.x + (h / 6.0) * (R1[i][j][0] + R2[i][j][0]);
                vn.y = vb.y + (h / 6.0) * (R1[i][j][1] + R2[i][j][1]);
                *reinterpret_cast<double2 *>(Ub + a * ldu + b) = un;
                *reinterpret_cast<double2 *>(Vb + a * ldu + b) = vn;
              }
              *reinterpret_cast<double2 *>(Us + a * ldu + (b ^ sw)) = un;  // U_{s+1}; after stage 4: U(t+h)
            }
          }
        }
        if (t < d) {
          if (separable) {
            if (s < 4) H[t * ldh + t] = hdv[s * dp + t];
          } else {
            accq += wgt * kq;
            accp += wgt * kp;
            if (s < 4) {
              qsa = qa + cnext * kq;
              psa = pa + cnext * kp;
              qs[t] = qsa;
            } else {
              qa += h / 6.0 * accq;
              pa += h / 6.0 * accp;
              q[t] = qa;
              p[t] = pa;
            }
          }
        }
        __syncthreads();
        PT(3);
        if (s < 4 && !separable) {
          vpart = pot_eval<TPT>(P, qs, g, H, ldh, scr, scr2, t, gid, P.type == POT_ROTATED_MORSE, true);
          if (P.type == POT_ROTATED_MORSE) rotated_hessian_mma<NW>(P, scr2, H, ldh, warp, lane);
        }
        PT(1);
      }
      if (separable && t < d) { accS = sacc[t]; e4 = se4[t]; }
      // ================= correlation partial sums (need only q, p): reduced by warps 0 and 1, consumed by
      // thread 0 after the LU, whose barriers order the shared-memory traffic =================
      {
        double v8[8];
        double v6[6];
        corr_terms<TPT>(E, q, p, E.zt + (size_t)traj * 2 * d, dqv, dpv, v6, t, gid);
        if (warp < 2) {
#pragma unroll
          for (int i = 0; i < 6; ++i) v8[i] = v6[i];
          v8[6] = accS;
          v8[7] = e4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v8[i] += __shfl_xor_sync(0xffffffffu, v8[i], o);
          }
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) red[i * 2 + warp] = v8[i];
          }
        }
      }
      // ================= prefactor: assembly into registers (transposed: det A^T = det A), LU =================
      double2 lo[MC], hi[MC];
      if (E.diag) {
        const double sb0 = (lane < d) ? E.sgi[lane] : 0.0, isb0 = (lane < d) ? E.isgi[lane] : 0.0;
        const double sb1 = (lane + 32 < d) ? E.sgi[lane + 32] : 0.0, isb1 = (lane + 32 < d) ? E.isgi[lane + 32] : 0.0;
#pragma unroll
        for (int m = 0; m < MC; ++m) {
          const int a = warp + NW * m;
          lo[m] = hi[m] = make_double2(0.0, 0.0);
          if (a < d) {
            const double sa = 0.5 * E.sgt[a], isa = 0.5 * E.isgt[a];
            const double *ur = Ub + a * ldu, *vr = Vb + a * ldu;
            if (lane < d)
              lo[m] = make_double2(sa * ur[lane] * isb0 + isa * vr[d + lane] * sb0,
                                   -sa * ur[d + lane] * sb0 + isa * vr[lane] * isb0);
            if (lane + 32 < d)
              hi[m] = make_double2(sa * ur[lane + 32] * isb1 + isa * vr[d + lane + 32] * sb1,
                                   -sa * ur[d + lane + 32] * sb1 + isa * vr[lane + 32] * isb1);
          }
        }
      } else {
        const int ldc = dr | 1;
        // staging of the factor matrices needs d ldr + dr ldl doubles of the stage-operand region (it does for every
        // supported d; the DFMA version is the fallback)
        if (DK * mma_frag_ld(dr) + dr * mma_frag_ld(d) <= DK * ldu)
          prefactor_assemble_mma<NW>(E, Ub, Vb, ldu, Cm, ldc, H, ldh, Us, t, warp, lane);
        else
          prefactor_assemble<TPT>(E, Ub, Vb, ldu, Cm, ldc, H, t, gid);
#pragma unroll
        for (int m = 0; m < MC; ++m) {
          const int a = warp + NW * m;
          lo[m] = hi[m] = make_double2(0.0, 0.0);
          if (a < dr) {
            if (lane < dr) lo[m] = Cm[a * ldc + lane];
            if (lane + 32 < dr) hi[m] = Cm[a * ldc + lane + 32];
          }
        }
      }
      PT(4);
      if (split) {
        double2 *mat = cm + ((size_t)step * ntw + tl) * dr * dr;
#pragma unroll
        for (int m = 0; m < MC; ++m) {
          const int a = warp + NW * m;
          if (a < dr) {
            if (lane < dr) mat[a * dr + lane] = lo[m];
            if (lane + 32 < dr) mat[a * dr + lane + 32] = hi[m];
          }
        }
        __syncthreads();                                      // red[] of warps 0 and 1 is complete
        if (t == 0) {                                         // only thread 0 carries the action
          S += h / 6.0 * (red[12] + red[13]);
          double *ax = aux + ((size_t)step * ntw + tl) * 8;
#pragma unroll
          for (int i = 0; i < 6; ++i) ax[i] = red[i * 2] + red[i * 2 + 1];
          ax[6] = S;
          ax[7] = red[14] + red[15];
        }
      }
      const double2 det = split ? make_double2(1.0, 0.0) : lu_det_regs<NW, MC, 0>(lo, hi, dr, lush, warp, lane);
      PT(5);
      if (t == 0 && !split) {
        double v6[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) v6[i] = red[i * 2] + red[i * 2 + 1];
        S += h / 6.0 * (red[12] + red[13]);
        sign = track_sign(sign, c2, det);
        c2 = det;
        cc = csqrt_principal(det);
        double2 ca, ki;
        corr_finish(E, v6, S, cc, sign, wvi, ca, ki);
        double *row = cacc + step * 5;
        row[0] += ca.x; row[1] += ca.y; row[2] += ki.x; row[3] += ki.y; row[4] += red[14] + red[15];
      }
      PT(6);
      if (!E.diag) {
        // the dense assembly used the Us region for the prefactor matrix and H as scratch: restore the stage-1
        // operand U(t+h) for the next step
        __syncthreads();
        if (step + 1 < nsteps) {
          for (int idx = t; idx < (DK - d) * ldu; idx += TPT) Us[d * ldu + idx] = 0.0;
          for (int idx = t; idx < d * d; idx += TPT) {
            const int a = idx / d, b = 2 * (idx % d);
            *reinterpret_cast<double2 *>(Us + a * ldu + (b ^ swz(a))) = *reinterpret_cast<const double2 *>(Ub + a * ldu + b);
          }
        }
        if (separable)
          for (int i = t; i < ((d + 7) & ~7) * ldh; i += TPT) H[i] = 0.0;
        __syncthreads();
      }
      PT(7);
    }
    // ---- write back
    if (t < d) { rec[t] = q[t]; rec[d + t] = p[t]; }
    if (t == 0) {
      rec[2 * d] = S;
      if (!split) {
        E.c2[traj] = c2;
        E.c[traj] = cc;
        E.sign[traj] = sign;
      }
    }
    for (int idx = t; idx < NE; idx += TPT) {
      const int a = idx / W, b = idx % W;
      rec[E.qps + idx] = Ub[a * ldu + b];
      rec[E.qps + NE + idx] = Vb[a * ldu + b];
    }
    __syncthreads();
    PT(8);
  }
  // per-CTA correlation sums of this launch (each row is written by exactly one CTA: no memset, no atomics)
  if (!split)
    for (int i = t; i < 5 * nsteps; i += TPT)
      partials[((size_t)gg * nsteps_total + step0 + i / 5) * 5 + i % 5] = cacc[i];
}

// ------------------------------------------------------------------ host-side dispatch -------
struct MmaConfig { int wm, wn, nwm, nwn; };

static bool mma_config(int d, MmaConfig &c) {
  if (d < 17 || d > 62) return false;
  if (d <= 32) { c = {2, 2, 2, 4}; return true; }        //  8 warps: rows <= 32, cols <= 64
  if (d <= 48) { c = {2, 3, 3, 4}; return true; }        // 12 warps: rows <= 48, cols <= 96
  if (d <= 60) { c = {2, 5, 4, 3}; return true; }        // 12 warps: rows <= 64, cols <= 120 (d = 60: no padding)
  c = {2, 4, 4, 4};                                       // 16 warps: rows <= 64, cols <= 128
  return true;
}

static bool mma_supported(int d) {
  MmaConfig c;
  return mma_config(d, c);
}

static void mma_leading_dims(int d, int &ldu, int &ldh) {
  ldu = 2 * d;
  while (ldu % 16 != 8) ldu += 2;      // 128-bit owner accesses conflict-free
  const int dk = (d + 3) & ~3;
  ldh = dk;
  while (ldh % 16 != 4 && ldh % 16 != 12) ldh += 4;   // conflict-free A-fragment loads
}

static int mma_threads(int d) {
  MmaConfig c;
  mma_config(d, c);
  return 32 * c.nwm * c.nwn;
}

// cm == nullptr: fused mode over the whole ensemble (launches of at most MMA_KMAX steps); else split mode: ONE launch of
// nsteps <= MMA_KMAX steps over the window [traj0, traj0 + ntw)
template <int WM, int WN, int NWM, int NWN, int MC>
static cudaError_t launch_mma_t(int grid, size_t smem, const EngDev &E, const PotDev &P, double h, int nsteps,
                                double *partials, const SmemLayout &L, cudaStream_t st, int traj0, int ntw, double2 *cm,
                                double *aux) {
  auto kern = k_hk_mma<WM, WN, NWM, NWN, MC>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce != cudaSuccess) return ce;
  for (int s0 = 0; s0 < nsteps; s0 += MMA_KMAX) {
    const int ns = (nsteps - s0 < MMA_KMAX) ? nsteps - s0 : MMA_KMAX;
    kern<<<grid, 32 * NWM * NWN, smem, st>>>(E, P, h, ns, s0, nsteps, partials, L, traj0, ntw, cm, aux);
    ce = cudaGetLastError();
    if (ce != cudaSuccess) return ce;
  }
  return cudaSuccess;
}

static cudaError_t launch_mma(int grid, int threads, size_t smem, const EngDev &E, const PotDev &P, double h,
                              int nsteps, double *partials, const SmemLayout &L, cudaStream_t st, int traj0 = 0, int ntw = -1,
                              double2 *cm = nullptr, double *aux = nullptr) {
  MmaConfig c;
  if (!mma_config(E.d, c) || threads != 32 * c.nwm * c.nwn) return cudaErrorInvalidValue;
  if (ntw < 0) ntw = E.n;
  // last parameter: LU columns per thread = ceil(max d of the bucket / warps)
  if (E.d <= 32) return launch_mma_t<2, 2, 2, 4, 4>(grid, smem, E, P, h, nsteps, partials, L, st, traj0, ntw, cm, aux);
  if (E.d <= 48) return launch_mma_t<2, 3, 3, 4, 4>(grid, smem, E, P, h, nsteps, partials, L, st, traj0, ntw, cm, aux);
  if (E.d <= 60) return launch_mma_t<2, 5, 4, 3, 5>(grid, smem, E, P, h, nsteps, partials, L, st, traj0, ntw, cm, aux);
  return launch_mma_t<2, 4, 4, 4, 4>(grid, smem, E, P, h, nsteps, partials, L, st, traj0, ntw, cm, aux);
}

}  // namespace sc
