// sc_small.cuh -- fused Herman-Kluk step kernel for small systems (d <= 12; any rank d' <= d): register-resident
// monodromy columns.
//
// k_hk_generic (sc_kernels.cuh) gives a whole warp to one trajectory and keeps its state in shared memory; at d = 5 that
// is 100 monodromy elements on 32 lanes, three shared-memory round trips per element and stage, and a warp-wide barrier
// between every phase.  Here one THREAD owns one real column of U = [Mqq|Mqp] and V = [Mpq|Mpp] (the 2 d columns of
// the monodromy equations are independent, propagators.py:342-357), so a trajectory takes 2 d lanes and a warp carries
// 32 / (2 d) trajectories (three at d = 5, sixteen at d = 1).  Per stage a thread does kv = -H us on its own column from
// registers; the Hessian (d doubles for the separable potentials, d x d otherwise) is the only thing that goes through
// shared memory.  Thread c < d also owns mode c of (q, p).  The prefactor matrix is assembled column by column (thread c
// < d' gets column c; the p-half columns arrive by shuffles for diagonal Gamma and through a small staging buffer for
// dense Gamma) and its LU runs column-distributed: the owner of column k finds the pivot in its registers and broadcasts
// the pivot row index and its column, every other column updates itself.
//
// Same reference semantics as k_hk_generic: RK4 bookkeeping (header of sc_kernels.cuh), prefactor propagators.py:959-1001,
// branch tracking propagators.py:1035-1051, contributions propagators.py:868-909.
#pragma once
#include "sc_kernels.cuh"

namespace sc {

template <int D, int DR>
struct SmallCfg {
  static constexpr int GS = 2 * D;                       // lanes per trajectory
  static constexpr int NGW = 32 / GS;                    // trajectories per warp
  static constexpr int WARPS = 4;
  static constexpr int LDH = (D + 1) & ~1;
  static constexpr int MIN_CTAS = D <= 5 ? 4 : 2;        // register budget: 168 / 255 per thread
  // per-CTA constant block (doubles): hess0 or Q | 1/m | sgt | isgt | L1 | L2 | R1 | R2 | otA | otB | otC
  static constexpr int O_IM = D * LDH, O_SGT = O_IM + LDH, O_ISGT = O_SGT + LDH, O_L1 = O_ISGT + LDH, O_L2 = O_L1 + DR * D,
                       O_R1 = O_L2 + DR * D, O_R2 = O_R1 + D * DR, O_OA = O_R2 + D * DR, O_OB = O_OA + D * D, O_OC = O_OB + D * D;
  static constexpr int O_PC = (O_OC + D * D + 1) & ~1;   // per-mode constants read once per step: q0 p0 wR wG oA oB oC sgi isgi
  static constexpr int CTA_CONST = O_PC + 9 * LDH;
  static constexpr int PER_GROUP = ((D * LDH + 8 * LDH + GS * 2 * DR + 8 * D + 10) + 1) & ~1;   // doubles
  static constexpr size_t SMEM = sizeof(double) * (size_t)(CTA_CONST + WARPS * NGW * PER_GROUP);
};

template <int N>
__device__ __forceinline__ double2 pick_row(const double2 (&col)[N], int p) {
  double2 r = col[0];
#pragma unroll
  for (int i = 1; i < N; ++i)
    if (p == i) r = col[i];
  return r;
}

// PT / DG: potential type and diagonal-width flag folded at compile time (-1: read from the arguments)
template <int D, int DR, int PT = -1, int DG = -1>
__global__ void __launch_bounds__(128, SmallCfg<D, DR>::MIN_CTAS)
k_hk_small(EngDev E, PotDev P, double h, int nsteps, double *partials) {
  using Cfg = SmallCfg<D, DR>;
  constexpr int GS = Cfg::GS, NGW = Cfg::NGW, LDH = Cfg::LDH, W = 2 * D, NE = 2 * D * D;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) double smem[];
  double *cmat = smem, *cim = smem + D * LDH;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / GS, c = lane - g * GS;
  const bool lane_ok = g < NGW;                           // lanes beyond the last whole group idle (they still take part in shuffles)
  const int gbase = lane_ok ? g * GS : 0;
  double *gs = smem + Cfg::CTA_CONST + (size_t)(warp * NGW + (lane_ok ? g : 0)) * Cfg::PER_GROUP;
  double *Hs = gs, *hdv = Hs + D * LDH, *qs = hdv + LDH, *scr = qs + LDH, *scr2 = scr + LDH, *dqv = scr2 + LDH,
         *dpv = dqv + LDH, *v4s = dpv + LDH, *v5s = v4s + LDH, *Tst = v5s + LDH, *red = Tst + GS * 2 * DR, *lead = red + 8 * D;
  // lead: state of the group's leader thread (S, sign, det, sqrt(det), initial overlap), kept out of everybody's registers
  const int ptype = PT >= 0 ? PT : P.type;
  const bool diag = DG >= 0 ? (DG != 0) : (E.diag != 0);
  const bool all_harmonic = (PT == POT_MORSE && DG == 1) ? false : (P.all_harmonic != 0);   // the compile-time Morse case is anharmonic
  const bool separable = ptype == POT_MORSE || ptype == POT_NONHARMONIC;
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
    const int a = i / D, k = i % D;
    cmat[a * LDH + k] = (ptype == POT_HARMONIC) ? P.hess0[i] : (ptype == POT_ROTATED_MORSE ? P.Q[i] : 0.0);
  }
  if (threadIdx.x < D) cim[threadIdx.x] = P.imass[threadIdx.x];
  {
    // the small constant tables go to shared memory once: addressed by immediates, no pointer registers in the step loop
    double *cst = smem;
    const int t = threadIdx.x, nt = blockDim.x;
    for (int i = t; i < D; i += nt) {
      double *pc = cst + Cfg::O_PC + i;
      pc[0] = E.q0[i]; pc[LDH] = E.p0[i]; pc[2 * LDH] = E.wR[i]; pc[3 * LDH] = E.wG[i];
      if (diag) {
        pc[4 * LDH] = E.otA[i]; pc[5 * LDH] = E.otB[i]; pc[6 * LDH] = E.otC[i]; pc[7 * LDH] = E.sgi[i]; pc[8 * LDH] = E.isgi[i];
      }
    }
    if (diag) {
      for (int i = t; i < D; i += nt) { cst[Cfg::O_SGT + i] = E.sgt[i]; cst[Cfg::O_ISGT + i] = E.isgt[i]; }
    } else {
      // rank d' <= DR: the factors are zero-padded to DR rows / columns and the padded diagonal of the prefactor matrix is
      // set to one below, which leaves the determinant unchanged
      const int dr = E.dr;
      for (int i = t; i < DR * D; i += nt) {
        const int ap = i / D;                                  // L1, L2: (d' x d) row-major
        cst[Cfg::O_L1 + i] = ap < dr ? E.L1[i] : 0.0;
        cst[Cfg::O_L2 + i] = ap < dr ? E.L2[i] : 0.0;
        const int b = i / DR, j = i % DR;                      // R1, R2: (d x d') row-major
        cst[Cfg::O_R1 + i] = j < dr ? E.R1[b * dr + j] : 0.0;
        cst[Cfg::O_R2 + i] = j < dr ? E.R2[b * dr + j] : 0.0;
      }
      for (int i = t; i < D * D; i += nt) { cst[Cfg::O_OA + i] = E.otA[i]; cst[Cfg::O_OB + i] = E.otB[i]; cst[Cfg::O_OC + i] = E.otC[i]; }
    }
  }
  __syncthreads();
  const double *cst = smem;
  const double *Hc = (ptype == POT_HARMONIC) ? cmat : Hs;
  const int wg = blockIdx.x * Cfg::WARPS + warp, NWG = gridDim.x * Cfg::WARPS;   // warp index, warps in the grid
  const bool modal = lane_ok && c < D;
  // per-mode constants of this thread
  double im_c = 0, pos0c = 0, grad0c = 0;
  double pa1 = 0, pa2 = 0, pa3 = 0;                       // potential parameters of mode c
  const int cm = c < D ? c : 0;
#define SC_PC(k) cst[Cfg::O_PC + (k) * LDH + cm]
  if (modal) {
    im_c = P.imass[c];
    if (ptype == POT_HARMONIC) { pos0c = P.pos0[c]; grad0c = P.grad0[c]; }
    if (ptype == POT_MORSE || ptype == POT_ROTATED_MORSE) {
      if (all_harmonic) pa1 = P.omega[c] * P.omega[c];
      else { pa1 = P.a[c]; pa2 = P.D[c]; }
    }
    if (ptype == POT_NONHARMONIC) { pa1 = P.eps[c]; pa2 = P.b[c]; }
  }
  (void)pa3;

  for (int t0 = 0; t0 < E.n; t0 += NWG * NGW) {
    const int traj = t0 + wg * NGW + g;
    const bool valid = lane_ok && traj < E.n;
    const bool mode_c = valid && c < D;
    double *rec = E.rec + (size_t)(valid ? traj : 0) * E.rs;
    double u[D], v[D];
#pragma unroll
    for (int a = 0; a < D; ++a) {
      u[a] = valid ? rec[E.qps + a * W + c] : 0.0;
      v[a] = valid ? rec[E.qps + NE + a * W + c] : 0.0;
    }
    double qa = 0, pa = 0;
    if (mode_c) {
      qa = rec[c]; pa = rec[D + c];
      const double *zt = E.zt + (size_t)traj * 2 * D;
      v4s[c] = (SC_PC(0) - zt[c]) * SC_PC(2);
      v5s[c] = (zt[D + c] - SC_PC(1)) * SC_PC(3);
    }
    if (valid && c == 0) {
      const double2 c2 = E.c2[traj], cc = E.c[traj], wvi = E.wvi[traj];
      lead[0] = rec[2 * D]; lead[1] = E.sign[traj]; lead[2] = c2.x; lead[3] = c2.y; lead[4] = cc.x; lead[5] = cc.y;
      lead[6] = wvi.x; lead[7] = wvi.y;
    }

    for (int step = 0; step < nsteps; ++step) {
      // ================= one classical RK4 step =================
      double us[D], R1[D], R2[D];
#pragma unroll
      for (int a = 0; a < D; ++a) { us[a] = u[a]; R1[a] = 0.0; R2[a] = 0.0; }
      double qsa = qa, psa = pa, accq = 0, accp = 0, accS = 0, e4 = 0;
      constexpr int STAGE_UNROLL = (PT >= 0) ? 4 : 1;     // the folded kernels are small enough to unroll the stages
#pragma unroll STAGE_UNROLL
      for (int s = 1; s <= 4; ++s) {
        // ---- potential at the stage point: gradient component gc of this mode, Hessian into shared memory
        double vpart = 0.0, gc = 0.0;
        __syncwarp();                                     // the previous stage has read the Hessian
        if (separable) {
          if (mode_c) {
            double hd;
            if (ptype == POT_MORSE) {
              if (all_harmonic) {
                vpart = 0.5 * pa1 * qsa * qsa; gc = pa1 * qsa; hd = pa1;
              } else {
                const double e = exp(-pa1 * qsa);
                vpart = pa2 * (1.0 - e) * (1.0 - e);
                gc = 2.0 * pa1 * pa2 * e * (1.0 - e);
                hd = 2.0 * pa1 * pa1 * pa2 * e * (2.0 * e - 1.0);
              }
            } else {
              const double e1 = exp(-pa2 * qsa), e2 = exp(-2.0 * pa2 * qsa);
              vpart = pa1 / (2.0 * pa2 * pa2) * (1.0 - e1) * (1.0 - e1) + (1.0 - pa1) * 0.5 * qsa * qsa;
              gc = pa1 / pa2 * (e1 - e2) + (1.0 - pa1) * qsa;
              hd = pa1 * (2.0 * e2 - e1) + (1.0 - pa1);
            }
            hdv[c] = hd;
            if (c == 0) vpart -= P.origin;
          }
        } else if (ptype == POT_HARMONIC) {
          if (mode_c) scr[c] = qsa - pos0c;
          __syncwarp();
          if (mode_c) {
            double hd = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) hd = fma(cmat[c * LDH + j], scr[j], hd);
            gc = grad0c + hd;
            vpart = scr[c] * grad0c + 0.5 * scr[c] * hd;
            if (c == 0) vpart += P.e0 - P.origin;
          }
        } else {                                          // rotated Morse: r = Q^T x, grad = Q g, H = Q diag(h) Q^T
          if (mode_c) qs[c] = qsa;
          __syncwarp();
          if (mode_c) {
            double r = 0.0;
#pragma unroll
            for (int i = 0; i < D; ++i) r = fma(cmat[i * LDH + c], qs[i], r);
            double gi, hi;
            if (all_harmonic) {
              vpart = 0.5 * pa1 * r * r; gi = pa1 * r; hi = pa1;
            } else {
              const double e = exp(-pa1 * r);
              vpart = pa2 * (1.0 - e) * (1.0 - e);
              gi = 2.0 * pa1 * pa2 * e * (1.0 - e);
              hi = 2.0 * pa1 * pa1 * pa2 * e * (2.0 * e - 1.0);
            }
            scr[c] = gi;
            scr2[c] = hi;
            if (c == 0) vpart -= P.origin;
          }
          __syncwarp();
          if (mode_c) {
#pragma unroll
            for (int k = 0; k < D; ++k) gc = fma(cmat[c * LDH + k], scr[k], gc);
          }
          if (valid) {
            for (int idx = c; idx < D * D; idx += GS) {
              const int i = idx / D, j = idx % D;
              double sacc = 0.0;
#pragma unroll
              for (int k = 0; k < D; ++k) sacc = fma(cmat[i * LDH + k] * scr2[k], cmat[j * LDH + k], sacc);
              Hs[i * LDH + j] = sacc;
            }
          }
        }
        __syncwarp();
        // ---- kv = -H us on this thread's column
        double kv[D];
        if (separable) {
#pragma unroll
          for (int a = 0; a < D; ++a) kv[a] = -hdv[a] * us[a];
        } else {
#pragma unroll
          for (int a = 0; a < D; ++a) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) acc = fma(Hc[a * LDH + k], us[k], acc);
            kv[a] = -acc;
          }
        }
        // ---- accumulators and next stage operand (header of sc_kernels.cuh)
        const double cnext = (s == 3) ? h : 0.5 * h;
        const double wgt = (s == 1 || s == 4) ? 1.0 : 2.0;
#pragma unroll
        for (int a = 0; a < D; ++a) {
          const double ima = cim[a];
          if (s == 1) {
            R1[a] = kv[a];
            R2[a] = 0.0;
            us[a] = u[a] + 0.5 * h * v[a] * ima;
          } else if (s == 2) {
            us[a] = u[a] + 0.5 * h * (v[a] + 0.5 * h * R1[a]) * ima;
            R1[a] += kv[a];
            R2[a] = kv[a];
          } else if (s == 3) {
            us[a] = u[a] + h * (v[a] + 0.5 * h * R2[a]) * ima;
            R1[a] += kv[a];
            R2[a] += kv[a];
          } else {
            R2[a] += kv[a];
            const double ub = u[a], vb = v[a];
            u[a] = ub + h * vb * ima + (h * h / 6.0) * R1[a] * ima;
            v[a] = vb + (h / 6.0) * (R1[a] + R2[a]);
          }
        }
        {
          const double kq = psa * im_c, kp = -gc;
          const double tk = 0.5 * psa * psa * im_c;
          accS += wgt * (tk - vpart);
          if (s == 4) e4 = tk + vpart;
          accq += wgt * kq;
          accp += wgt * kp;
          if (s < 4) {
            qsa = qa + cnext * kq;
            psa = pa + cnext * kp;
          } else {
            qa += h / 6.0 * accq;
            pa += h / 6.0 * accp;
          }
        }
      }
      // ================= prefactor matrix, column c on thread c < DR =================
      double2 Cc[DR];
      if (diag) {
        // diagonal width matrices (d' = d): element-wise scaling; the p-half columns live D lanes further up
#pragma unroll
        for (int a = 0; a < DR; ++a) {
          const int aa = a < D ? a : 0;
          const double up = __shfl_sync(FULL, u[aa], (lane + D) & 31);    // Mqp[a][c]
          const double vp = __shfl_sync(FULL, v[aa], (lane + D) & 31);    // Mpp[a][c]
          const double sa = cst[Cfg::O_SGT + aa], isa = cst[Cfg::O_ISGT + aa];
          const double sgi_c = SC_PC(7), isgi_c = SC_PC(8);
          Cc[a] = make_double2(0.5 * (sa * u[aa] * isgi_c + isa * vp * sgi_c), 0.5 * (-sa * up * sgi_c + isa * v[aa] * isgi_c));
        }
      } else {
        // dense: left factors on the own column, then the right factors over the staged columns
        __syncwarp();
        if (valid) {
#pragma unroll
          for (int ap = 0; ap < DR; ++ap) {
            double t1 = 0.0, t2 = 0.0;
#pragma unroll
            for (int a = 0; a < D; ++a) {
              t1 = fma(cst[Cfg::O_L1 + ap * D + a], u[a], t1);
              t2 = fma(cst[Cfg::O_L2 + ap * D + a], v[a], t2);
            }
            Tst[c * 2 * DR + ap] = t1;
            Tst[c * 2 * DR + DR + ap] = t2;
          }
        }
        __syncwarp();
        const int cr = c < DR ? c : 0;
#pragma unroll
        for (int ap = 0; ap < DR; ++ap) Cc[ap] = make_double2(0.0, 0.0);
#pragma unroll
        for (int b = 0; b < D; ++b) {
          const double r1 = cst[Cfg::O_R1 + b * DR + cr], r2 = cst[Cfg::O_R2 + b * DR + cr];
          const double *Tq = Tst + b * 2 * DR, *Tp = Tst + (D + b) * 2 * DR;
#pragma unroll
          for (int ap = 0; ap < DR; ++ap) {
            Cc[ap].x = fma(Tq[ap], r1, fma(Tp[DR + ap], r2, Cc[ap].x));     // L1 Mqq R1 + L2 Mpp R2
            Cc[ap].y = fma(Tq[DR + ap], r1, fma(-Tp[ap], r2, Cc[ap].y));    // L2 Mpq R1 - L1 Mqp R2
          }
        }
#pragma unroll
        for (int ap = 0; ap < DR; ++ap) {
          Cc[ap].x *= 0.5;
          Cc[ap].y *= 0.5;
          if (ap >= E.dr && ap == c) Cc[ap].x = 1.0;             // identity on the padded diagonal (d' < DR)
        }
      }
      // ================= determinant: column-distributed LU, implicit partial pivoting =================
      double2 det = make_double2(1.0, 0.0);
      {
        unsigned done = 0u;
        int inversions = 0;
#pragma unroll
        for (int k = 0; k < DR; ++k) {
          const int owner = (gbase + k) & 31;
          int p = 0;
          double best = -1.0;
#pragma unroll
          for (int i = 0; i < DR; ++i) {
            const double m = Cc[i].x * Cc[i].x + Cc[i].y * Cc[i].y;
            if (!((done >> i) & 1u) && m > best) { best = m; p = i; }
          }
          p = __shfl_sync(FULL, p, owner);
          const double2 cp = pick_row<DR>(Cc, p);                          // pivot-row element of the own column
          const double2 pv = make_double2(__shfl_sync(FULL, cp.x, owner), __shfl_sync(FULL, cp.y, owner));
          det = cmul(det, pv);
          inversions += __popc(done >> p);                                 // earlier pivots with a larger row index
          done |= 1u << p;
          if (k + 1 < DR) {
            const double2 ip = cinv(pv);
#pragma unroll
            for (int i = 0; i < DR; ++i) {
              const double2 ck = make_double2(__shfl_sync(FULL, Cc[i].x, owner), __shfl_sync(FULL, Cc[i].y, owner));
              if (!((done >> i) & 1u)) {
                const double2 f = cmul(ck, ip);
                Cc[i].x -= f.x * cp.x - f.y * cp.y;
                Cc[i].y -= f.x * cp.y + f.y * cp.x;
              }
            }
          }
        }
        if (inversions & 1) { det.x = -det.x; det.y = -det.y; }
      }
      // ================= correlation contributions =================
      {
        const double p0c = SC_PC(1);
        const double dq = SC_PC(0) - qa, dp = p0c - pa;
        double v0, v1;
        if (diag) {
          v0 = -0.5 * (dq * SC_PC(4) * dq + dp * SC_PC(5) * dp);
          v1 = -p0c * dq + dq * SC_PC(6) * dp;
        } else {
          if (mode_c) { dqv[c] = dq; dpv[c] = dp; }
          __syncwarp();
          double sa = 0.0, sb = 0.0, sc_ = 0.0;
          const int cm_ = c < D ? c : 0;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            sa = fma(cst[Cfg::O_OA + j * D + cm_], dqv[j], sa);
            sb = fma(cst[Cfg::O_OB + j * D + cm_], dpv[j], sb);
            sc_ = fma(cst[Cfg::O_OC + j * D + cm_], dqv[j], sc_);
          }
          v0 = -0.5 * (dq * sa + dp * sb);
          v1 = -p0c * dq + dp * sc_;
        }
        if (mode_c) {
          red[0 * D + c] = v0; red[1 * D + c] = v1; red[2 * D + c] = dq * SC_PC(2); red[3 * D + c] = -dp * SC_PC(3);
          red[6 * D + c] = accS; red[7 * D + c] = e4;
        }
      }
      __syncwarp();
      double row5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      if (valid && c == 0) {
        double v8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double *src = (i == 4) ? v4s : (i == 5) ? v5s : red + i * D;
          double sacc = 0.0;
#pragma unroll
          for (int a = 0; a < D; ++a) sacc += src[a];
          v8[i] = sacc;
        }
        const double S = lead[0] + h / 6.0 * v8[6];
        const double sign = track_sign(lead[1], make_double2(lead[2], lead[3]), det);
        const double2 cc = csqrt_principal(det);
        lead[0] = S; lead[1] = sign; lead[2] = det.x; lead[3] = det.y; lead[4] = cc.x; lead[5] = cc.y;
        double2 ca, ki;
        const double v6[6] = {v8[0], v8[1], v8[2], v8[3], v8[4], v8[5]};
        corr_finish(E, v6, S, cc, sign, make_double2(lead[6], lead[7]), ca, ki);
        row5[0] = ca.x; row5[1] = ca.y; row5[2] = ki.x; row5[3] = ki.y; row5[4] = v8[7];
      }
      if (NGW > 1) {
        // one partial row per warp: fixed-order tree over the lanes (non-leaders carry zeros)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) row5[i] += __shfl_xor_sync(FULL, row5[i], o);
        }
      }
      if (lane == 0) {
        double *row = partials + ((size_t)wg * nsteps + step) * 5;
#pragma unroll
        for (int i = 0; i < 5; ++i) row[i] += row5[i];
      }
      if (E.snap != nullptr && valid) {
        // snapshot of the new time for the fused Walton-Manolopoulos launch (sc_wm.cuh: k_wm_fused)
        const size_t item = (size_t)step * E.n + traj;
        double *sr = E.snap + item * E.rs;
#pragma unroll
        for (int a = 0; a < D; ++a) {
          sr[E.qps + a * W + c] = u[a];
          sr[E.qps + NE + a * W + c] = v[a];
        }
        if (c < D) { sr[c] = qa; sr[D + c] = pa; }
        if (c == 0) {
          sr[2 * D] = lead[0];
          E.snap_c[item] = make_double2(lead[4], lead[5]);
          E.snap_sign[item] = lead[1];
        }
      }
    }
    // ---- write back
    if (valid) {
#pragma unroll
      for (int a = 0; a < D; ++a) {
        rec[E.qps + a * W + c] = u[a];
        rec[E.qps + NE + a * W + c] = v[a];
      }
      if (c < D) { rec[c] = qa; rec[D + c] = pa; }
      if (c == 0) {
        rec[2 * D] = lead[0];
        E.c2[traj] = make_double2(lead[2], lead[3]);
        E.c[traj] = make_double2(lead[4], lead[5]);
        E.sign[traj] = lead[1];
      }
    }
    __syncwarp();
  }
#undef SC_PC
}

inline bool small_supported(const EngDev &E, const PotDev &P) {
  if (!(P.type == POT_MORSE || P.type == POT_NONHARMONIC || P.type == POT_HARMONIC || P.type == POT_ROTATED_MORSE)) return false;
  if (E.diag && E.dr != E.d) return false;
  const int d = E.d;
  return d >= 1 && d <= 12;                                     // any rank d' <= d (zero-padded factors)
}

template <int D, int DR, int PT = -1, int DG = -1>
static cudaError_t launch_small_t(int sm_count, const EngDev &E, const PotDev &P, double h, int nsteps, double *partials,
                                  int &nrows_groups, bool plan_only, cudaStream_t st) {
  using Cfg = SmallCfg<D, DR>;
  auto kern = k_hk_small<D, DR, PT, DG>;
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (ce != cudaSuccess) return ce;
    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, Cfg::SMEM);
    if (ce != cudaSuccess) return ce;
    if (per_sm < 1) per_sm = 1;
  }
  const int per_cta = Cfg::WARPS * Cfg::NGW;
  int grid = (E.n + per_cta - 1) / per_cta;
  if (grid > sm_count * per_sm) grid = sm_count * per_sm;
  if (grid < 1) grid = 1;
  nrows_groups = grid * Cfg::WARPS;
  if (plan_only) return cudaSuccess;
  kern<<<grid, 128, Cfg::SMEM, st>>>(E, P, h, nsteps, partials);
  return cudaGetLastError();
}

// plan_only: returns the number of partial rows per step (warps in the grid) without launching
inline cudaError_t launch_small(int sm_count, const EngDev &E, const PotDev &P, double h, int nsteps, double *partials,
                                int &nrows_groups, bool plan_only, cudaStream_t st) {
#define SC_SMALL_CASE(D_, R_) \
  if (E.d == D_ && E.dr == R_) return launch_small_t<D_, R_>(sm_count, E, P, h, nsteps, partials, nrows_groups, plan_only, st)
  // the two small BASELINE configs with everything folded: C1 / C2 (anharmonic Morse, diagonal widths), C3 (harmonic, dense widths)
  if (E.d == 5 && E.dr == 5 && P.type == POT_MORSE && E.diag && !P.all_harmonic)
    return launch_small_t<5, 5, POT_MORSE, 1>(sm_count, E, P, h, nsteps, partials, nrows_groups, plan_only, st);
  if (E.d == 12 && E.dr == 6 && P.type == POT_HARMONIC && !E.diag)
    return launch_small_t<12, 6, POT_HARMONIC, 0>(sm_count, E, P, h, nsteps, partials, nrows_groups, plan_only, st);
  SC_SMALL_CASE(12, 6);
#undef SC_SMALL_CASE
#define SC_SMALL_CASE(D_) \
  if (E.d == D_) return launch_small_t<D_, D_>(sm_count, E, P, h, nsteps, partials, nrows_groups, plan_only, st)
  SC_SMALL_CASE(1); SC_SMALL_CASE(2); SC_SMALL_CASE(3); SC_SMALL_CASE(4); SC_SMALL_CASE(5); SC_SMALL_CASE(6);
  SC_SMALL_CASE(7); SC_SMALL_CASE(8); SC_SMALL_CASE(9); SC_SMALL_CASE(10); SC_SMALL_CASE(11); SC_SMALL_CASE(12);
#undef SC_SMALL_CASE
  return cudaErrorInvalidValue;
}

}  // namespace sc
