// sc_wm.cuh -- Walton-Manolopoulos (Filinov-smoothed) prefactor pieces and per-trajectory contributions to the
// autocorrelation / IC correlation functions (reference: propagators.py:1132-1389, 1577-1719).
//
// One group (a warp for d <= 8, a 128-thread CTA above) owns one trajectory; all intermediate matrices live in
// shared memory.  The HK kernel has already advanced (q, p, M, S) and the HK prefactor C for this step; this
// kernel rebuilds the WM matrices from the monodromy blocks, inverts A' (2d' x 2d') and M' (d' x d') by
// Gauss-Jordan elimination with partial pivoting (the determinants are the products of the scaled pivots),
// tracks the sqrt branches of det A and det M and reduces the contributions.
//
// The per-trajectory routine is __host__ __device__: tests/emul compiles it for the host with one "thread"
// per group so that the formulas can be checked on machines without a GPU.
//
// Note on eqn (55): as coded in the reference b0 = gradL - i (Mqz^T P - Eqz^T p) with gradL = i (Mqz^T P - Eqz^T p)
// (propagators.py:1167-1180, 1262-1264), i.e. b0 vanishes identically; the terms it multiplies (the b0 parts of
// eqns 60 and 74) are therefore dropped here: pi_t = P, pi_i = p, eps = -1/2 (p0-p) iGi0 (p0-p).
#pragma once
#include "sc_device.cuh"

#if defined(__CUDA_ARCH__)
#define SC_LDG(p) __ldg(p)
#else
#define SC_LDG(p) (*(p))
#endif
#define SC_HD __host__ __device__ __forceinline__

namespace sc {

enum WMMode { WM_STEP = 0, WM_INIT = 1, WM_CORR = 2, WM_DIAG = 3 };   // DIAG: pieces of the wavefunction diagnostics only

struct WMDev {
  int d, dr;
  const double *G0, *Gi, *Gt, *iGi0, *iG0;  // width matrices and pseudo-inverses (d x d)
  const double *GiG;                        // Gamma_0 iGi0
  const double *Cqq;                        // Gamma_0 - Gamma_0 iGi0 Gamma_0   (eqn 69, trajectory independent)
  const double *U;                          // d x dr
  const double *q0, *p0;
  const double *n1;                         // -tau1/m of the potential in use (d)
  double alpha, beta;
  double pref;                              // sqrt(detG0) detGt^(1/4) detGi^(1/4) / sqrt(detGi0)   (eqn 85)
  double2 *prevA, *prevM;                   // branch trackers "detA", "detM" (n)
  double *signA, *signM;
  const double *winv;                       // 1 / (probi (2 pi)^d)
  // wavefunction diagnostics (coefficients / wavefunction / norm, propagators.py:1391-1575): trajectory-minor arrays
  // [component][trajectory] written in mode WM_DIAG
  double pref_coef, diag_inv_norm;          // detG0^(1/4) detGt^(1/4) detGi^(1/4) / sqrt(detGi0);  1 / ntraj
  double2 *dg_v;                            // (n) coefficients, eqn (75)
  double2 *dg_CQQ, *dg_UC, *dg_CP;          // CQQ (d^2), U^T CQQ (dr d), U^T CQQ U (dr^2)
  double2 *dg_D, *dg_DP;                    // dvec = CqQ^T (q0 - q) + i PIQ (d), U^T dvec (dr)
  double *dg_Q;                             // current positions (d)
};

// shared-memory workspace of one group, offsets in double2 units
struct WMLayout {
  int mq, vr, bq, a, t1, ap, iap, gt, gti, cQQ, cqQ, x1, x2, im, rqq, rQQ, rqQ, mp, imp, vc, total;
};

__host__ __device__ inline WMLayout make_wm_layout(int d, int dr) {
  WMLayout L;
  const int D2 = 2 * d, R2 = 2 * dr, d2 = d * d;
  int o = 0;
  if (dr == d) {
    // full-rank widths: no projections (wm_trajectory), A is inverted in place of A', and the arrays that are born after
    // Gt~ / Gti (eqns 57, 59) reuse the space of those that are dead by then -- A (destroyed by the inversion), A^-1 and BQ
    // (last read by eqns 57, 59).  494 instead of 894 complex numbers at d = 5: four CTAs of four trajectories per SM
    // instead of three, and 0.77 MB instead of 1.8 MB per slab at d = 60
    L.mq = o; o += d * D2;
    L.vr = o; o += 4 * d;
    L.bq = o; o += d * D2;
    L.a = o; o += D2 * D2;
    L.t1 = o; o += d * D2;
    L.ap = L.a;
    L.iap = o; o += R2 * R2;
    L.gt = o; o += d2;
    L.gti = o; o += d2;
    L.x1 = L.a; L.cqQ = L.a + d2; L.cQQ = L.a + 2 * d2; L.x2 = L.a + 3 * d2;
    L.rqq = L.iap; L.rQQ = L.iap + d2; L.rqQ = L.iap + 2 * d2; L.mp = L.iap + 3 * d2;
    L.imp = L.bq; L.im = L.bq + d2;
    L.vc = o; o += 12 * d + R2 + 4;
    L.total = o;
    return L;
  }
  L.mq = o; o += d * D2;            // [Mqq|Mqp], [Mpq|Mpp] as two real d x 2d arrays
  L.vr = o; o += 4 * d;             // real vectors: Q, P, qi, pi, dq, dQ, v2, PIq
  L.bq = o; o += d * D2;
  L.a = o; o += D2 * D2;
  L.t1 = o; o += D2 * D2;
  L.ap = o; o += R2 * R2;
  L.iap = o; o += R2 * R2;
  L.gt = o; o += d2;
  L.gti = o; o += d2;
  L.cQQ = o; o += d2;
  L.cqQ = o; o += d2;
  L.x1 = o; o += d2;
  L.x2 = o; o += d2;
  L.im = o; o += d2;
  L.rqq = o; o += d2;
  L.rQQ = o; o += d2;
  L.rqQ = o; o += d2;
  L.mp = o; o += dr * dr;
  L.imp = o; o += dr * dr;
  L.vc = o; o += 12 * d + R2 + 4;   // complex vectors: PIQ, v2c, Pq, PQ, w, u1..u6, pivot column, scalars
  L.total = o;
  return L;
}

template <int TPT>
SC_HD void gsync(int gid) {
#if defined(__CUDA_ARCH__)
  Group<TPT>::sync(gid);
#else
  (void)gid;
#endif
}

SC_HD double2 c_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SC_HD double2 c_inv(double2 a) {
  if (fabs(a.x) >= fabs(a.y)) {
    const double r = a.y / a.x, den = a.x + a.y * r;
    return make_double2(1.0 / den, -r / den);
  }
  const double r = a.x / a.y, den = a.x * r + a.y;
  return make_double2(r / den, -1.0 / den);
}
SC_HD double2 c_sqrt(double2 z) {
  const double r = hypot(z.x, z.y);
  if (r == 0.0) return make_double2(0.0, 0.0);
  if (z.x >= 0.0) {
    const double sr = sqrt(0.5 * (r + z.x));
    return make_double2(sr, z.y / (2.0 * sr));
  }
  const double si = copysign(sqrt(0.5 * (r - z.x)), z.y);
  return make_double2(z.y / (2.0 * si), si);
}
SC_HD double2 c_exp(double2 z) {
  const double e = exp(z.x);
  return make_double2(e * cos(z.y), e * sin(z.y));
}

// Gauss-Jordan inverse with partial pivoting of the n x n complex matrix a (destroyed) into inv; returns
// prod_k (pivot_k * pivscale) with the sign of the row permutation, i.e. det(a * pivscale).
// col: n complex entries of scratch.  Every thread of the group returns the determinant.
template <int TPT>
SC_HD double2 gj_inverse(double2 *a, double2 *inv, int n, double pivscale, double2 *col, int t, int gid) {
  for (int idx = t; idx < n * n; idx += TPT) inv[idx] = make_double2((idx / n == idx % n) ? 1.0 : 0.0, 0.0);
  double2 det = make_double2(1.0, 0.0);
  gsync<TPT>(gid);
  for (int k = 0; k < n; ++k) {
    // pivot search (redundantly by every thread: n <= 2 d' entries, broadcast reads); first maximum wins
    int p = k;
    double best = -1.0;
    for (int i = k; i < n; ++i) {
      const double2 v = a[i * n + k];
      const double m = v.x * v.x + v.y * v.y;
      if (m > best) { best = m; p = i; }
    }
    gsync<TPT>(gid);
    if (p != k) {
      for (int j = t; j < 2 * n; j += TPT) {
        double2 *m = (j < n) ? a : inv;
        const int jj = (j < n) ? j : j - n;
        const double2 tmp = m[k * n + jj];
        m[k * n + jj] = m[p * n + jj];
        m[p * n + jj] = tmp;
      }
      det.x = -det.x;
      det.y = -det.y;
      gsync<TPT>(gid);
    }
    const double2 piv = a[k * n + k];
    det = c_mul(det, make_double2(piv.x * pivscale, piv.y * pivscale));
    const double2 ipiv = c_inv(piv);
    gsync<TPT>(gid);
    for (int j = t; j < 2 * n; j += TPT) {
      double2 *m = (j < n) ? a : inv;
      const int jj = (j < n) ? j : j - n;
      m[k * n + jj] = c_mul(m[k * n + jj], ipiv);
    }
    for (int i = t; i < n; i += TPT) col[i] = a[i * n + k];
    gsync<TPT>(gid);
    for (int idx = t; idx < n * 2 * n; idx += TPT) {
      const int i = idx / (2 * n), j = idx % (2 * n);
      if (i == k) continue;
      double2 *m = (j < n) ? a : inv;
      const int jj = (j < n) ? j : j - n;
      const double2 f = col[i], u = m[k * n + jj];
      double2 v = m[i * n + jj];
      v.x -= f.x * u.x - f.y * u.y;
      v.y -= f.x * u.y + f.y * u.x;
      m[i * n + jj] = v;
    }
    gsync<TPT>(gid);
  }
  return det;
}

SC_HD double wm_track(double sign, double2 zprev, double2 z) {
  return (zprev.x < 0.0 && z.x < 0.0 && zprev.y * z.y < 0.0) ? -sign : sign;
}

// All WM work of one trajectory.  acc4 (thread 0 of the group only): running sums of the contributions to
// C_auto (re, im) and k_ic (re, im), weights 1/(probi (2 pi)^d) applied, 1/N not applied.
// DC > 0: d = dr = DC at compile time (full-rank widths): the index arithmetic (idx / D2, idx % R2, ...) and the loop bounds of the
// ~30 small matrix products fold into constants -- the runtime-d kernel spends most of its issue slots on them (14.9 k warp
// instructions per trajectory-step at d = 5, profiles/ncu_r02_k_wm_fused.txt)
template <int TPT, int DC = 0>
SC_HD void wm_trajectory(const EngDev &E, const WMDev &W, const WMLayout &L, double2 *ws, int traj, int mode, int t,
                         int gid, double *acc4) {
  const int d = DC > 0 ? DC : W.d, dr = DC > 0 ? DC : W.dr, D2 = 2 * d, R2 = 2 * dr, d2 = d * d;
  double *MQ = reinterpret_cast<double *>(ws + L.mq), *MP = MQ + d * D2;
  double *Qv = reinterpret_cast<double *>(ws + L.vr), *Pv = Qv + d, *qi = Pv + d, *pi = qi + d;
  double *dq = pi + d, *dQ = dq + d, *v2 = dQ + d, *PIq = v2 + d;
  double2 *BQ = ws + L.bq, *A = ws + L.a, *T1 = ws + L.t1, *AP = ws + L.ap, *IAP = ws + L.iap;
  double2 *GT = ws + L.gt, *GTI = ws + L.gti, *CQQ = ws + L.cQQ, *CqQ = ws + L.cqQ, *X1 = ws + L.x1, *X2 = ws + L.x2;
  double2 *IM = ws + L.im, *Rqq = ws + L.rqq, *RQQ = ws + L.rQQ, *RqQ = ws + L.rqQ, *MPr = ws + L.mp, *IMP = ws + L.imp;
  double2 *PIQ = ws + L.vc, *v2c = PIQ + d, *Pq = v2c + d, *PQ = Pq + d, *wv = PQ + d, *u1 = wv + d, *u2 = u1 + d,
          *u3 = u2 + d, *u4 = u3 + d, *u5 = u4 + d, *u6 = u5 + d, *col = u6 + d;
  const double *rec = E.rec + (size_t)traj * E.rs;
  const double *zt = E.zt + (size_t)traj * 2 * d;

  for (int idx = t; idx < d * D2; idx += TPT) {
    MQ[idx] = rec[E.qps + idx];
    MP[idx] = rec[E.qps + 2 * d2 + idx];
  }
  for (int i = t; i < d; i += TPT) {
    Qv[i] = rec[i];
    Pv[i] = rec[d + i];
    qi[i] = zt[i];
    pi[i] = zt[d + i];
    dq[i] = SC_LDG(W.q0 + i) - zt[i];
    dQ[i] = SC_LDG(W.q0 + i) - rec[i];
    v2[i] = SC_LDG(W.p0 + i) - zt[d + i];          // p0 - pi_i  (pi_i = p, see the note on b0)
  }
  const double S = rec[2 * d];
  gsync<TPT>(gid);

  // eqn (53): BQ = Gamma_t Mqz + i Mpz   (d x 2d)
  for (int idx = t; idx < d * D2; idx += TPT) {
    const int i = idx / D2, k = idx % D2;
    double s = 0.0;
    for (int j = 0; j < d; ++j) s += SC_LDG(W.Gt + i * d + j) * MQ[j * D2 + k];
    BQ[idx] = make_double2(s, MP[idx]);
  }
  gsync<TPT>(gid);
  // eqn (50): A = 2 F - hessL + Mqz^T Gt Mqz + Eqz^T Gi Eqz + 2i (Mpz^T Mqz - Epz^T Eqz)
  for (int idx = t; idx < D2 * D2; idx += TPT) {
    const int i = idx / D2, l = idx % D2;
    double f = 0.0, gi = 0.0;
    if (i < d && l < d) { f = W.alpha * SC_LDG(W.G0 + i * d + l); gi = SC_LDG(W.Gi + i * d + l); }
    if (i >= d && l >= d) f = W.beta * SC_LDG(W.iG0 + (i - d) * d + (l - d));
    double hs = 0.0, mg = 0.0, pm = 0.0;
    for (int j = 0; j < d; ++j) {
      const double mqi = MQ[j * D2 + i], mpi = MP[j * D2 + i], mql = MQ[j * D2 + l], mpl = MP[j * D2 + l];
      // hessL blocks (propagators.py:1184-1187): rows i < d use Mpq^T [Mqq|Mqp], rows i >= d use Mqp^T [Mpq|Mpp]
      hs += (i < d) ? mpi * mql : mqi * mpl;
      pm += mpi * mql;
      mg += mqi * BQ[j * D2 + l].x;
    }
    const double ee = (i >= d && l == i - d) ? 1.0 : 0.0;
    A[idx] = make_double2(2.0 * f + mg + gi, -hs + 2.0 * (pm - ee));
  }
  gsync<TPT>(gid);
  // full-rank widths: U is orthogonal, the projection A' = U2^T A U2 is a similarity transform (same determinant, and
  // U2 A'^-1 U2^T = A^-1), so the four projection products are skipped and A is inverted in place of A'
  const bool fullrank = (dr == d);
  // A' = U2^T A U2 with U2 = blockdiag(U, U)
  if (!fullrank)
  for (int idx = t; idx < R2 * D2; idx += TPT) {
    const int a = idx / D2, j = idx % D2;
    const int off = (a < dr) ? 0 : d, aa = (a < dr) ? a : a - dr;
    double2 s = make_double2(0.0, 0.0);
    for (int i = 0; i < d; ++i) {
      const double u = SC_LDG(W.U + i * dr + aa);
      const double2 v = A[(off + i) * D2 + j];
      s.x += u * v.x;
      s.y += u * v.y;
    }
    T1[idx] = s;
  }
  if (!fullrank) gsync<TPT>(gid);
  if (!fullrank)
  for (int idx = t; idx < R2 * R2; idx += TPT) {
    const int a = idx / R2, b = idx % R2;
    const int off = (b < dr) ? 0 : d, bb = (b < dr) ? b : b - dr;
    double2 s = make_double2(0.0, 0.0);
    for (int i = 0; i < d; ++i) {
      const double u = SC_LDG(W.U + i * dr + bb);
      const double2 v = T1[a * D2 + off + i];
      s.x += u * v.x;
      s.y += u * v.y;
    }
    AP[idx] = s;
  }
  if (!fullrank) gsync<TPT>(gid);
  // A'^-1 and det(A' / (2 sqrt(alpha beta)))   (propagators.py:1255, 1328-1332)
  const double2 detA = gj_inverse<TPT>(fullrank ? A : AP, IAP, R2, 1.0 / (2.0 * sqrt(W.alpha * W.beta)), col, t, gid);
  const double2 *Ainv = fullrank ? IAP : A;
  // A^-1 = U2 A'^-1 U2^T  -> A
  if (!fullrank)
  for (int idx = t; idx < D2 * R2; idx += TPT) {
    const int i = idx / R2, b = idx % R2;
    const int off = (i < d) ? 0 : dr, ii = (i < d) ? i : i - d;
    double2 s = make_double2(0.0, 0.0);
    for (int a = 0; a < dr; ++a) {
      const double u = SC_LDG(W.U + ii * dr + a);
      const double2 v = IAP[(off + a) * R2 + b];
      s.x += u * v.x;
      s.y += u * v.y;
    }
    T1[idx] = s;
  }
  if (!fullrank) gsync<TPT>(gid);
  if (!fullrank)
  for (int idx = t; idx < D2 * D2; idx += TPT) {
    const int i = idx / D2, j = idx % D2;
    const int off = (j < d) ? 0 : dr, jj = (j < d) ? j : j - d;
    double2 s = make_double2(0.0, 0.0);
    for (int b = 0; b < dr; ++b) {
      const double u = SC_LDG(W.U + jj * dr + b);
      const double2 v = T1[i * R2 + off + b];
      s.x += u * v.x;
      s.y += u * v.y;
    }
    A[idx] = s;
  }
  if (!fullrank) gsync<TPT>(gid);
  // T1 = BQ A^-1  (d x 2d)
  for (int idx = t; idx < d * D2; idx += TPT) {
    const int i = idx / D2, k = idx % D2;
    double2 s = make_double2(0.0, 0.0);
    for (int j = 0; j < D2; ++j) {
      const double2 b = BQ[i * D2 + j], v = Ainv[j * D2 + k];
      s.x += b.x * v.x - b.y * v.y;
      s.y += b.x * v.y + b.y * v.x;
    }
    T1[idx] = s;
  }
  gsync<TPT>(gid);
  // eqn (57): Gt~ = Gamma_t - BQ A^-1 BQ^T ; eqn (59): Gti = BQ A^-1 Bq^T with Bq = [Gamma_i | -i 1]
  for (int idx = t; idx < d2; idx += TPT) {
    const int i = idx / d, l = idx % d;
    double2 s = make_double2(0.0, 0.0), g = make_double2(0.0, 0.0);
    for (int k = 0; k < D2; ++k) {
      const double2 a = T1[i * D2 + k], b = BQ[l * D2 + k];
      s.x += a.x * b.x - a.y * b.y;
      s.y += a.x * b.y + a.y * b.x;
    }
    for (int k = 0; k < d; ++k) {
      const double2 a = T1[i * D2 + k];
      const double gi = SC_LDG(W.Gi + l * d + k);
      g.x += a.x * gi;
      g.y += a.y * gi;
    }
    const double2 a = T1[i * D2 + d + l];   // times -i
    g.x += a.y;
    g.y -= a.x;
    GT[idx] = make_double2(SC_LDG(W.Gt + idx) - s.x, -s.y);
    GTI[idx] = g;
  }
  gsync<TPT>(gid);
  // X1 = Gti iGi0 ; eqn (71): CqQ = Gamma_0 iGi0 Gti^T
  for (int idx = t; idx < d2; idx += TPT) {
    const int i = idx / d, k = idx % d;
    double2 s = make_double2(0.0, 0.0), c = make_double2(0.0, 0.0);
    for (int j = 0; j < d; ++j) {
      const double2 g = GTI[i * d + j];
      const double ig = SC_LDG(W.iGi0 + j * d + k);
      s.x += g.x * ig;
      s.y += g.y * ig;
      const double gg = SC_LDG(W.GiG + i * d + j);
      const double2 h = GTI[k * d + j];
      c.x += gg * h.x;
      c.y += gg * h.y;
    }
    X1[idx] = s;
    CqQ[idx] = c;
  }
  gsync<TPT>(gid);
  // eqn (70): CQQ = Gt~ - Gti iGi0 Gti^T ; eqns (72, 73): PIq, PIQ
  for (int idx = t; idx < d2; idx += TPT) {
    const int i = idx / d, l = idx % d;
    double2 s = make_double2(0.0, 0.0);
    for (int k = 0; k < d; ++k) {
      const double2 a = X1[i * d + k], b = GTI[l * d + k];
      s.x += a.x * b.x - a.y * b.y;
      s.y += a.x * b.y + a.y * b.x;
    }
    CQQ[idx] = make_double2(GT[idx].x - s.x, GT[idx].y - s.y);
  }
  for (int i = t; i < d; i += TPT) {
    double s1 = 0.0;
    double2 s2 = make_double2(0.0, 0.0);
    for (int k = 0; k < d; ++k) {
      s1 += SC_LDG(W.GiG + i * d + k) * v2[k];
      s2.x += X1[i * d + k].x * v2[k];
      s2.y += X1[i * d + k].y * v2[k];
    }
    PIq[i] = SC_LDG(W.p0 + i) - s1;
    PIQ[i] = make_double2(Pv[i] + s2.x, s2.y);
    v2c[i] = make_double2(PIQ[i].x - SC_LDG(W.p0 + i), PIQ[i].y);      // PIQ - p0
  }
  gsync<TPT>(gid);
  if (mode == WM_DIAG) {
    // pieces of coefficients() / wavefunction() / norm() (propagators.py:1391-1575): nothing of the trackers is touched
    const int n = E.n;
    for (int idx = t; idx < dr * d; idx += TPT) {                 // UC = U^T CQQ
      const int a = idx / d, j = idx % d;
      double2 sum = make_double2(0.0, 0.0);
      for (int i = 0; i < d; ++i) {
        const double u = SC_LDG(W.U + i * dr + a);
        sum.x += u * CQQ[i * d + j].x;
        sum.y += u * CQQ[i * d + j].y;
      }
      T1[idx] = sum;
    }
    for (int i = t; i < d; i += TPT) {                            // dvec = CqQ^T (q0 - q) + i PIQ   (propagators.py:1511)
      double2 sum = make_double2(-PIQ[i].y, PIQ[i].x);
      for (int b = 0; b < d; ++b) {
        sum.x += CqQ[b * d + i].x * dq[b];
        sum.y += CqQ[b * d + i].y * dq[b];
      }
      u1[i] = sum;
    }
    gsync<TPT>(gid);
    for (int idx = t; idx < dr * dr; idx += TPT) {                // CP = U^T CQQ U
      const int a = idx / dr, b = idx % dr;
      double2 sum = make_double2(0.0, 0.0);
      for (int j = 0; j < d; ++j) {
        const double u = SC_LDG(W.U + j * dr + b);
        sum.x += T1[a * d + j].x * u;
        sum.y += T1[a * d + j].y * u;
      }
      W.dg_CP[(size_t)idx * n + traj] = sum;
    }
    for (int a = t; a < dr; a += TPT) {
      double2 sum = make_double2(0.0, 0.0);
      for (int i = 0; i < d; ++i) {
        const double u = SC_LDG(W.U + i * dr + a);
        sum.x += u * u1[i].x;
        sum.y += u * u1[i].y;
      }
      W.dg_DP[(size_t)a * n + traj] = sum;
    }
    for (int idx = t; idx < d2; idx += TPT) W.dg_CQQ[(size_t)idx * n + traj] = CQQ[idx];
    for (int idx = t; idx < dr * d; idx += TPT) W.dg_UC[(size_t)idx * n + traj] = T1[idx];
    for (int i = t; i < d; i += TPT) {
      W.dg_D[(size_t)i * n + traj] = u1[i];
      W.dg_Q[(size_t)i * n + traj] = Qv[i];
    }
    if (t == 0) {
      // eqn (75) as coded in coefficients(), propagators.py:1407-1430; eps of eqn (74) with b0 = 0
      double viv = 0.0, qCq = 0.0, piq = 0.0;
      for (int i = 0; i < d; ++i) {
        double s1 = 0.0, s2 = 0.0;
        for (int k = 0; k < d; ++k) {
          s1 += SC_LDG(W.iGi0 + i * d + k) * v2[k];
          s2 += SC_LDG(W.Cqq + i * d + k) * dq[k];
        }
        viv += v2[i] * s1;
        qCq += dq[i] * s2;
        piq += PIq[i] * dq[i];
      }
      const double2 c = E.c[traj];
      const double s0 = E.sign[traj] * W.signA[traj] * W.pref_coef * W.winv[traj] * W.diag_inv_norm;
      double2 v = c_mul(make_double2(s0 * c.x, s0 * c.y), make_double2(cos(S), sin(S)));
      v = c_mul(v, c_inv(c_sqrt(detA)));
      v = c_mul(v, c_exp(make_double2(-0.5 * viv - 0.5 * qCq, -piq)));
      W.dg_v[traj] = v;
    }
    gsync<TPT>(gid);
    return;
  }
  // eqn (78): M = Gamma_0 + CQQ, projected: M' = U^T M U   (full rank: M itself, see above)
  if (fullrank)
    for (int idx = t; idx < d2; idx += TPT) MPr[idx] = make_double2(SC_LDG(W.G0 + idx) + CQQ[idx].x, CQQ[idx].y);
  if (!fullrank)
  for (int idx = t; idx < dr * d; idx += TPT) {
    const int a = idx / d, j = idx % d;
    double2 s = make_double2(0.0, 0.0);
    for (int i = 0; i < d; ++i) {
      const double u = SC_LDG(W.U + i * dr + a);
      s.x += u * (SC_LDG(W.G0 + i * d + j) + CQQ[i * d + j].x);
      s.y += u * CQQ[i * d + j].y;
    }
    T1[idx] = s;
  }
  if (!fullrank) gsync<TPT>(gid);
  if (!fullrank)
  for (int idx = t; idx < dr * dr; idx += TPT) {
    const int a = idx / dr, b = idx % dr;
    double2 s = make_double2(0.0, 0.0);
    for (int j = 0; j < d; ++j) {
      const double u = SC_LDG(W.U + j * dr + b);
      s.x += T1[a * d + j].x * u;
      s.y += T1[a * d + j].y * u;
    }
    MPr[idx] = s;
  }
  gsync<TPT>(gid);
  const double2 detM = gj_inverse<TPT>(MPr, IMP, dr, 1.0 / (2.0 * M_PI), col, t, gid);
  const double2 *IMi = fullrank ? IMP : IM;
  // M^-1 = U M'^-1 U^T
  if (!fullrank)
  for (int idx = t; idx < d * dr; idx += TPT) {
    const int i = idx / dr, b = idx % dr;
    double2 s = make_double2(0.0, 0.0);
    for (int a = 0; a < dr; ++a) {
      const double u = SC_LDG(W.U + i * dr + a);
      s.x += u * IMP[a * dr + b].x;
      s.y += u * IMP[a * dr + b].y;
    }
    T1[idx] = s;
  }
  if (!fullrank) gsync<TPT>(gid);
  if (!fullrank)
  for (int idx = t; idx < d2; idx += TPT) {
    const int i = idx / d, j = idx % d;
    double2 s = make_double2(0.0, 0.0);
    for (int b = 0; b < dr; ++b) {
      const double u = SC_LDG(W.U + j * dr + b);
      s.x += T1[i * dr + b].x * u;
      s.y += T1[i * dr + b].y * u;
    }
    IM[idx] = s;
  }
  if (!fullrank) gsync<TPT>(gid);
  // X1 = CqQ M^-1 ; X2 = Gamma_0 M^-1
  for (int idx = t; idx < d2; idx += TPT) {
    const int i = idx / d, k = idx % d;
    double2 s = make_double2(0.0, 0.0), g = make_double2(0.0, 0.0);
    for (int j = 0; j < d; ++j) {
      const double2 c = CqQ[i * d + j], m = IMi[j * d + k];
      s.x += c.x * m.x - c.y * m.y;
      s.y += c.x * m.y + c.y * m.x;
      const double g0 = SC_LDG(W.G0 + i * d + j);
      g.x += g0 * m.x;
      g.y += g0 * m.y;
    }
    X1[idx] = s;
    X2[idx] = g;
  }
  gsync<TPT>(gid);
  // eqns (79-81): Rqq = Cqq - CqQ M^-1 CqQ^T ; RQQ = Gamma_0 - Gamma_0 M^-1 Gamma_0 ; RqQ = CqQ M^-1 Gamma_0
  for (int idx = t; idx < d2; idx += TPT) {
    const int i = idx / d, l = idx % d;
    double2 s = make_double2(0.0, 0.0), r = make_double2(0.0, 0.0), q = make_double2(0.0, 0.0);
    for (int k = 0; k < d; ++k) {
      const double2 a = X1[i * d + k], b = CqQ[l * d + k];
      s.x += a.x * b.x - a.y * b.y;
      s.y += a.x * b.y + a.y * b.x;
      const double g0 = SC_LDG(W.G0 + k * d + l);
      r.x += a.x * g0;
      r.y += a.y * g0;
      q.x += X2[i * d + k].x * g0;
      q.y += X2[i * d + k].y * g0;
    }
    Rqq[idx] = make_double2(SC_LDG(W.Cqq + idx) - s.x, -s.y);
    RqQ[idx] = r;
    RQQ[idx] = make_double2(SC_LDG(W.G0 + idx) - q.x, -q.y);
  }
  // eqns (82, 83): Pq, PQ ; w = M^-1 (PIQ - p0) for eqn (84)
  for (int i = t; i < d; i += TPT) {
    double2 s1 = make_double2(0.0, 0.0), s2 = s1, s3 = s1;
    for (int k = 0; k < d; ++k) {
      const double2 v = v2c[k];
      s1.x += X1[i * d + k].x * v.x - X1[i * d + k].y * v.y;
      s1.y += X1[i * d + k].x * v.y + X1[i * d + k].y * v.x;
      s2.x += X2[i * d + k].x * v.x - X2[i * d + k].y * v.y;
      s2.y += X2[i * d + k].x * v.y + X2[i * d + k].y * v.x;
      s3.x += IMi[i * d + k].x * v.x - IMi[i * d + k].y * v.y;
      s3.y += IMi[i * d + k].x * v.y + IMi[i * d + k].y * v.x;
    }
    Pq[i] = make_double2(PIq[i] - s1.x, -s1.y);
    PQ[i] = make_double2(SC_LDG(W.p0 + i) + s2.x, s2.y);
    wv[i] = s3;
  }
  gsync<TPT>(gid);

  // branch trackers (propagators.py:1035-1051)
  double sA = 1.0, sM = 1.0;
  if (mode == WM_STEP) {
    sA = wm_track(W.signA[traj], W.prevA[traj], detA);
    sM = wm_track(W.signM[traj], W.prevM[traj], detM);
  } else if (mode == WM_CORR) {
    sA = W.signA[traj];
    sM = W.signM[traj];
  }
  gsync<TPT>(gid);
  if (t == 0 && mode != WM_CORR) {
    W.signA[traj] = sA;
    W.signM[traj] = sM;
    W.prevA[traj] = detA;
    W.prevM[traj] = detM;
  }
  if (mode == WM_INIT) return;

  // matrix-vector products for the quadratic forms of eqns (85) and (100)
  for (int i = t; i < d; i += TPT) {
    double2 a1 = make_double2(0.0, 0.0), a2 = a1, a3 = a1, a4 = a1, a5 = a1, a6 = a1;
    for (int k = 0; k < d; ++k) {
      const double n1k = SC_LDG(W.n1 + k), dqk = dq[k], dQk = dQ[k];
      a1.x += Rqq[i * d + k].x * dqk; a1.y += Rqq[i * d + k].y * dqk;
      a2.x += Rqq[i * d + k].x * n1k; a2.y += Rqq[i * d + k].y * n1k;
      a3.x += RQQ[i * d + k].x * dQk; a3.y += RQQ[i * d + k].y * dQk;
      a4.x += RQQ[i * d + k].x * n1k; a4.y += RQQ[i * d + k].y * n1k;
      a5.x += RqQ[i * d + k].x * dQk; a5.y += RqQ[i * d + k].y * dQk;
      a6.x += RqQ[i * d + k].x * n1k; a6.y += RqQ[i * d + k].y * n1k;
    }
    u1[i] = a1; u2[i] = a2; u3[i] = a3; u4[i] = a4; u5[i] = a5; u6[i] = a6;
  }
  gsync<TPT>(gid);
  if (t == 0) {
    double2 e1 = make_double2(0.0, 0.0), e2 = e1, e3 = e1, dqRn = e1, dQRn = e1, dqRqQn = e1, nRn = e1, nRdQ = e1;
    double2 pq = e1, pQ = e1, Pqn = e1, PQn = e1, vMv = e1;
    double viv = 0.0;
    for (int i = 0; i < d; ++i) {
      const double n1i = SC_LDG(W.n1 + i), dqi = dq[i], dQi = dQ[i];
      e1.x += dqi * u1[i].x; e1.y += dqi * u1[i].y;         // dq Rqq dq
      dqRn.x += dqi * u2[i].x; dqRn.y += dqi * u2[i].y;     // dq Rqq n1
      e2.x += dQi * u3[i].x; e2.y += dQi * u3[i].y;         // dQ RQQ dQ
      dQRn.x += dQi * u4[i].x; dQRn.y += dQi * u4[i].y;     // dQ RQQ n1
      e3.x += dqi * u5[i].x; e3.y += dqi * u5[i].y;         // dq RqQ dQ
      dqRqQn.x += dqi * u6[i].x; dqRqQn.y += dqi * u6[i].y; // dq RqQ n1
      nRn.x += n1i * u6[i].x; nRn.y += n1i * u6[i].y;       // n1 RqQ n1
      nRdQ.x += n1i * u5[i].x; nRdQ.y += n1i * u5[i].y;     // n1 RqQ dQ
      pq.x += Pq[i].x * dqi; pq.y += Pq[i].y * dqi;
      pQ.x += PQ[i].x * dQi; pQ.y += PQ[i].y * dQi;
      Pqn.x += Pq[i].x * n1i; Pqn.y += Pq[i].y * n1i;
      PQn.x += PQ[i].x * n1i; PQn.y += PQ[i].y * n1i;
      const double2 m = c_mul(v2c[i], wv[i]);
      vMv.x += m.x; vMv.y += m.y;
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += SC_LDG(W.iGi0 + i * d + k) * v2[k];
      viv += v2[i] * s;
    }
    // eqn (74) with b0 = 0, eqn (84)
    const double2 gamma = make_double2(-0.5 * viv - 0.5 * vMv.x, -0.5 * vMv.y);
    // exponent of eqn (85): gamma - 1/2 e1 - 1/2 e2 + e3 - i pq + i pQ
    const double2 expo = make_double2(gamma.x - 0.5 * e1.x - 0.5 * e2.x + e3.x + pq.y - pQ.y,
                                      gamma.y - 0.5 * e1.y - 0.5 * e2.y + e3.y - pq.x + pQ.x);
    const double2 c = E.c[traj];
    const double s0 = E.sign[traj];
    double2 pref = make_double2(W.pref * s0 * c.x, W.pref * s0 * c.y);
    pref = c_mul(pref, make_double2(cos(S), sin(S)));
    double2 r = c_inv(c_sqrt(detA));
    pref = c_mul(pref, make_double2(sA * r.x, sA * r.y));
    r = c_inv(c_sqrt(detM));
    pref = c_mul(pref, make_double2(sM * r.x, sM * r.y));
    double2 cq = c_mul(pref, c_exp(expo));
    const double w = W.winv[traj];
    cq.x *= w;
    cq.y *= w;
    // eqn (100)
    const double2 nacQ = make_double2(dQRn.x - dqRqQn.x + PQn.y, dQRn.y - dqRqQn.y - PQn.x);
    const double2 nacq = make_double2(dqRn.x - nRdQ.x - Pqn.y, dqRn.y - nRdQ.y + Pqn.x);
    double2 k = c_mul(nacQ, nacq);
    k.x += nRn.x;
    k.y += nRn.y;
    k = c_mul(k, cq);
    acc4[0] += cq.x; acc4[1] += cq.y; acc4[2] += k.x; acc4[3] += k.y;
  }
  gsync<TPT>(gid);
}

#if defined(__CUDACC__)
template <int TPT>
__global__ void __launch_bounds__(TPT <= 128 ? 128 : TPT) k_wm(EngDev E, WMDev W, WMLayout L, int mode, double *partials) {
  extern __shared__ __align__(16) double2 wm_smem[];
  const int G = blockDim.x / TPT, gid = threadIdx.x / TPT, t = threadIdx.x % TPT;
  const int gg = blockIdx.x * G + gid, NG = gridDim.x * G;
  double2 *ws = wm_smem + (size_t)gid * L.total;
  double acc4[4] = {0.0, 0.0, 0.0, 0.0};
  for (int traj = gg; traj < E.n; traj += NG) wm_trajectory<TPT>(E, W, L, ws, traj, mode, t, gid, acc4);
  if (t == 0 && mode != WM_INIT) {
    double *row = partials + (size_t)gg * 4;
    row[0] = acc4[0]; row[1] = acc4[1]; row[2] = acc4[2]; row[3] = acc4[3];
  }
}

// d beyond the shared-memory envelope (about 29 modes; 21 for rank-deficient widths): the workspace of the CTA's trajectory lives in global memory (one slab
// per CTA, mostly L1/L2 hits); same code, one group of TPT threads per CTA
template <int TPT>
__global__ void __launch_bounds__(TPT) k_wm_global(EngDev E, WMDev W, WMLayout L, int mode, double *partials, double2 *gws) {
  const int t = threadIdx.x;
  double2 *ws = gws + (size_t)blockIdx.x * L.total;
  double acc4[4] = {0.0, 0.0, 0.0, 0.0};
  for (int traj = blockIdx.x; traj < E.n; traj += gridDim.x) wm_trajectory<TPT>(E, W, L, ws, traj, mode, t, 0, acc4);
  if (t == 0 && mode != WM_INIT) {
    double *row = partials + (size_t)blockIdx.x * 4;
    row[0] = acc4[0]; row[1] = acc4[1]; row[2] = acc4[2]; row[3] = acc4[3];
  }
}

// K time steps per launch: the HK kernel stored the record, sqrt(det) and sign of every (step, trajectory); a group walks the
// steps of its trajectory IN TIME ORDER (the detA / detM branch trackers are sequential) and adds the contributions of step k
// to its own row k of partials (ngroups, K, 4) -- zeroed by the host, one writer per row
template <int TPT, int DC = 0>
__global__ void __launch_bounds__(128, (DC == 5 ? 4 : 1)) k_wm_fused(EngDev E, WMDev W, WMLayout L, int nsteps, double *partials) {
  extern __shared__ __align__(16) double2 wm_smem[];
  const int G = blockDim.x / TPT, gid = threadIdx.x / TPT, t = threadIdx.x % TPT;
  const int gg = blockIdx.x * G + gid, NG = gridDim.x * G;
  double2 *ws = wm_smem + (size_t)gid * L.total;
  for (int traj = gg; traj < E.n; traj += NG) {
    for (int step = 0; step < nsteps; ++step) {
      EngDev Es = E;
      Es.rec = E.snap + (size_t)step * E.n * E.rs;
      Es.c = E.snap_c + (size_t)step * E.n;
      Es.sign = E.snap_sign + (size_t)step * E.n;
      double acc4[4] = {0.0, 0.0, 0.0, 0.0};
      wm_trajectory<TPT, DC>(Es, W, L, ws, traj, WM_STEP, t, gid, acc4);
      if (t == 0) {
        double *row = partials + ((size_t)gg * nsteps + step) * 4;
        row[0] += acc4[0]; row[1] += acc4[1]; row[2] += acc4[2]; row[3] += acc4[3];
      }
    }
  }
}

// out (K, 5): columns 0..3 = inv_norm * sum over groups of partials (ngroups, K, 4); column 4 = energy_rows[k][4]
__global__ void k_wm_reduce_k(const double *partials, int ngroups, int nsteps, double inv_norm, const double *energy_rows, double *out) {
  const int k = blockIdx.x, j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (j < 4) {
    double s = 0.0;
    for (int g = lane; g < ngroups; g += 32) s += partials[((size_t)g * nsteps + k) * 4 + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[k * 5 + j] = s * inv_norm;
  } else if (j == 4 && lane == 0) {
    out[k * 5 + 4] = energy_rows[k * 5 + 4];
  }
}

// ------------------------------------------------------------------ WM wavefunction diagnostics
constexpr int WMD_MAX = 16;      // largest d of the diagnostic kernels (per-thread local matrices)

// psi(x_k) = sum_n v_n exp(-1/2 y CQQ_n y + dvec_n . y), y = x_k - Q_n   (propagators.py:1434-1482); one CTA per grid point
__global__ void __launch_bounds__(256)
k_wm_wavefunction(int d, int n, int nx, const double *__restrict__ x, WMDev W, double2 *__restrict__ phi) {
  __shared__ double red[2][8];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (int k = blockIdx.x; k < nx; k += gridDim.x) {
    double sx = 0.0, sy = 0.0;
    for (int traj = t; traj < n; traj += 256) {
      double y[WMD_MAX];
      for (int a = 0; a < d; ++a) y[a] = x[(size_t)a * nx + k] - W.dg_Q[(size_t)a * n + traj];
      double2 ex = make_double2(0.0, 0.0);
      for (int a = 0; a < d; ++a) {
        double2 r = W.dg_D[(size_t)a * n + traj];
        for (int b = 0; b < d; ++b) {
          const double2 c = W.dg_CQQ[(size_t)(a * d + b) * n + traj];
          r.x -= 0.5 * c.x * y[b];
          r.y -= 0.5 * c.y * y[b];
        }
        ex.x += y[a] * r.x;
        ex.y += y[a] * r.y;
      }
      const double2 term = c_mul(W.dg_v[traj], c_exp(ex));
      sx += term.x;
      sy += term.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
    __syncthreads();
    if (lane == 0) { red[0][w] = sx; red[1][w] = sy; }
    __syncthreads();
    if (t == 0) {
      double ax = 0.0, ay = 0.0;
      for (int i = 0; i < 8; ++i) { ax += red[0][i]; ay += red[1][i]; }
      phi[k] = make_double2(ax, ay);
    }
  }
}

// |psi|^2 = sum_ij conj(v_i) O_ij v_j with the "overlap" of propagators.py:1520-1565: D = conj(CQQ_i) + CQQ_j in the non-zero
// subspace (dr x dr), its inverse applied to b = CQQ_j dQ + conj(d_i) + d_j and its determinant by Gaussian elimination with
// partial pivoting in per-thread local memory.  One CTA per bra i, threads over the kets j (trajectory-minor arrays: coalesced);
// partial sums (gridDim.x, 2), one writer per row.
__global__ void __launch_bounds__(128)
k_wm_norm(int d, int dr, int n, WMDev W, double *__restrict__ partials) {
  __shared__ double2 sCP[WMD_MAX * WMD_MAX], sDP[WMD_MAX];
  __shared__ double sQ[WMD_MAX], red[2][4];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  double accx = 0.0, accy = 0.0;
  const double inv2pi = 1.0 / (2.0 * M_PI);
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    __syncthreads();
    for (int idx = t; idx < dr * dr; idx += 128) sCP[idx] = W.dg_CP[(size_t)idx * n + i];
    if (t < dr) sDP[t] = W.dg_DP[(size_t)t * n + i];
    if (t < d) sQ[t] = W.dg_Q[(size_t)t * n + i];
    __syncthreads();
    const double2 vi = W.dg_v[i];
    for (int j = t; j < n; j += 128) {
      double dQ[WMD_MAX];
      for (int a = 0; a < d; ++a) dQ[a] = W.dg_Q[(size_t)a * n + j] - sQ[a];
      // -1/2 dQ CQQ_j dQ - d_j . dQ
      double2 ex = make_double2(0.0, 0.0);
      for (int a = 0; a < d; ++a) {
        double2 r = W.dg_D[(size_t)a * n + j];
        r.x = -r.x; r.y = -r.y;
        for (int b = 0; b < d; ++b) {
          const double2 c = W.dg_CQQ[(size_t)(a * d + b) * n + j];
          r.x -= 0.5 * c.x * dQ[b];
          r.y -= 0.5 * c.y * dQ[b];
        }
        ex.x += dQ[a] * r.x;
        ex.y += dQ[a] * r.y;
      }
      // b' = U^T (CQQ_j dQ + conj(d_i) + d_j),  D' = conj(CP_i) + CP_j
      double2 D[WMD_MAX * WMD_MAX], bp[WMD_MAX], z[WMD_MAX];
      for (int a = 0; a < dr; ++a) {
        double2 r = W.dg_DP[(size_t)a * n + j];
        r.x += sDP[a].x;
        r.y -= sDP[a].y;
        for (int b = 0; b < d; ++b) {
          const double2 c = W.dg_UC[(size_t)(a * d + b) * n + j];
          r.x += c.x * dQ[b];
          r.y += c.y * dQ[b];
        }
        bp[a] = r;
        z[a] = r;
        for (int b = 0; b < dr; ++b) {
          const double2 c = W.dg_CP[(size_t)(a * dr + b) * n + j];
          D[a * dr + b] = make_double2(sCP[a * dr + b].x + c.x, -sCP[a * dr + b].y + c.y);
        }
      }
      // Gaussian elimination with partial pivoting: det(D' / (2 pi)), z = D'^-1 b'
      double2 det = make_double2(1.0, 0.0);
      for (int k = 0; k < dr; ++k) {
        int p = k;
        double best = D[k * dr + k].x * D[k * dr + k].x + D[k * dr + k].y * D[k * dr + k].y;
        for (int r = k + 1; r < dr; ++r) {
          const double m = D[r * dr + k].x * D[r * dr + k].x + D[r * dr + k].y * D[r * dr + k].y;
          if (m > best) { best = m; p = r; }
        }
        if (p != k) {
          for (int c = k; c < dr; ++c) { const double2 tmp = D[k * dr + c]; D[k * dr + c] = D[p * dr + c]; D[p * dr + c] = tmp; }
          const double2 tmp = z[k]; z[k] = z[p]; z[p] = tmp;
          det.x = -det.x; det.y = -det.y;
        }
        const double2 piv = D[k * dr + k];
        det = c_mul(det, make_double2(piv.x * inv2pi, piv.y * inv2pi));
        const double2 ip = c_inv(piv);
        for (int r = k + 1; r < dr; ++r) {
          const double2 f = c_mul(D[r * dr + k], ip);
          for (int c = k + 1; c < dr; ++c) {
            const double2 u = D[k * dr + c];
            D[r * dr + c].x -= f.x * u.x - f.y * u.y;
            D[r * dr + c].y -= f.x * u.y + f.y * u.x;
          }
          z[r].x -= f.x * z[k].x - f.y * z[k].y;
          z[r].y -= f.x * z[k].y + f.y * z[k].x;
        }
      }
      for (int k = dr - 1; k >= 0; --k) {
        double2 r = z[k];
        for (int c = k + 1; c < dr; ++c) {
          const double2 u = D[k * dr + c];
          r.x -= u.x * z[c].x - u.y * z[c].y;
          r.y -= u.x * z[c].y + u.y * z[c].x;
        }
        z[k] = c_mul(r, c_inv(D[k * dr + k]));
      }
      for (int a = 0; a < dr; ++a) {                           // + 1/2 b'^T D'^-1 b'
        ex.x += 0.5 * (bp[a].x * z[a].x - bp[a].y * z[a].y);
        ex.y += 0.5 * (bp[a].x * z[a].y + bp[a].y * z[a].x);
      }
      double2 ol = c_mul(c_inv(c_sqrt(det)), c_exp(ex));
      ol = c_mul(ol, W.dg_v[j]);
      accx += vi.x * ol.x + vi.y * ol.y;                       // conj(v_i) * ...
      accy += vi.x * ol.y - vi.y * ol.x;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { accx += __shfl_xor_sync(0xffffffffu, accx, o); accy += __shfl_xor_sync(0xffffffffu, accy, o); }
  if (lane == 0) { red[0][w] = accx; red[1][w] = accy; }
  __syncthreads();
  if (t == 0) {
    partials[(size_t)blockIdx.x * 2] = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    partials[(size_t)blockIdx.x * 2 + 1] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
  }
}

__global__ void k_wm_norm_reduce(const double *partials, int nrows, double *out2) {
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (j >= 2) return;
  double s = 0.0;
  for (int g = lane; g < nrows; g += 32) s += partials[(size_t)g * 2 + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out2[j] = s;
}

// deterministic second pass: out[0..3] = inv_norm * sum over groups; out[4] = energy passthrough
__global__ void k_wm_reduce(const double *partials, int ngroups, double inv_norm, const double *energy_src, double *out) {
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (j < 4) {
    double s = 0.0;
    for (int g = lane; g < ngroups; g += 32) s += partials[(size_t)g * 4 + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[j] = s * inv_norm;
  } else if (j == 4 && lane == 0 && energy_src) {
    out[4] = energy_src[4];
  }
}

__global__ void k_wm_winv(const double *probi, double inv2pid, int n, double *winv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) winv[i] = inv2pid / probi[i];
}
#endif

}  // namespace sc
