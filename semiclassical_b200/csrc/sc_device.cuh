// sc_device.cuh -- device-side building blocks of the fused Herman-Kluk step kernels (sm_100a).
//
// A "group" is the set of threads that owns one trajectory: one warp (TPT == 32, several trajectories per
// CTA, small d) or the whole CTA (TPT == blockDim.x, large d).  All per-trajectory state lives in shared
// memory / registers for the duration of a launch (K fused time steps).
//
// Reference semantics restated here (file:line relative to /root/reference/semiclassical):
//   pot_eval_*       potentials.py:63-134, 265-327, 581-593
//   prefactor        propagators.py:959-1001   (assembly + U projection + determinant + principal sqrt)
//   branch tracking  propagators.py:1035-1051
//   overlap / IC     propagators.py:230-237, 868-909
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace sc {

enum PotType { POT_MORSE = 0, POT_HARMONIC = 2, POT_NONHARMONIC = 3, POT_ROTATED_MORSE = 4, POT_GDML = 5 };

struct PotDev {
  int type, d;
  const double *imass;  // 1/m (d)
  const double *n1;     // -tau1/m (d)  (hbar = 1, constant NAC vector)
  const double *omega, *a, *D;
  int all_harmonic;
  const double *pos0, *grad0, *hess0;
  double e0;            // energy0 (harmonic) ; c (gdml)
  const double *eps, *b;
  const double *Q;
  double origin;
  // sGDML
  int n_atoms, n_train, n_desc;
  const double *xs_train, *jx_alphas;
  double sig, gstd;
};

struct EngDev {
  int d, dr, n, diag;
  const double *L1, *L2, *R1, *R2;          // dense prefactor factors (dr x d, d x dr)
  const double *sgt, *isgt, *sgi, *isgi;    // diagonal-Gamma fast path (d)
  const double *otA, *otB, *otC;            // overlap <.,Gt|.,G0>: dense (d x d) or diagonal (d)
  double ot_fac;
  const double *q0, *p0;
  const double *wR, *wG;                    // R n1, (G0 iGi0)^T n1 for the potential in use
  double p0n1;
  double *rec;                              // trajectory-major records
  int rs, qps;                              // record stride, offset of U inside a record
  const double *zt;                         // (n, 2d) initial phase-space points
  const double2 *wvi;                       // (n) <qi,pi|phi0> / (probi (2 pi)^d)
  double2 *c2, *c;                          // (n) det and principal sqrt
  double *sign;                             // (n) branch sign of sqrt(det)
  // optional per-step snapshots (K-step fused Walton-Manolopoulos launches): record, sqrt(det), sign of (step, traj)
  double *snap;                             // (K, n, rs) or null
  double2 *snap_c;                          // (K, n)
  double *snap_sign;                        // (K, n)
};

// ------------------------------------------------------------------ complex helpers ---------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cinv(double2 a) {
  // Smith's algorithm
  if (fabs(a.x) >= fabs(a.y)) {
    const double r = a.y / a.x, den = a.x + a.y * r;
    return make_double2(1.0 / den, -r / den);
  } else {
    const double r = a.x / a.y, den = a.x * r + a.y;
    return make_double2(r / den, -1.0 / den);
  }
}
__device__ __forceinline__ double2 csqrt_principal(double2 z) {
  const double r = hypot(z.x, z.y);
  if (r == 0.0) return make_double2(0.0, 0.0);
  if (z.x >= 0.0) {
    const double sr = sqrt(0.5 * (r + z.x));
    return make_double2(sr, z.y / (2.0 * sr));
  } else {
    const double si = copysign(sqrt(0.5 * (r - z.x)), z.y);
    return make_double2(z.y / (2.0 * si), si);
  }
}
__device__ __forceinline__ double2 cexp(double re, double im) {
  double s, c;
  sincos(im, &s, &c);
  const double e = exp(re);
  return make_double2(e * c, e * s);
}

// ------------------------------------------------------------------ group synchronisation ---
template <int TPT>
struct Group {
  // bar id 0 is __syncthreads; groups inside a CTA use ids 1..15
  static __device__ __forceinline__ void sync(int gid) {
    if (TPT == 32) {
      __syncwarp();
    } else {
      (void)gid;
      __syncthreads();
    }
  }
};

// sum N values over the group; result returned to every thread. red: smem scratch (>= N * 32 doubles per group)
template <int TPT, int N>
__device__ __forceinline__ void group_reduce(double (&v)[N], double *red, int t, int gid) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  if (TPT == 32) return;
  const int w = t >> 5, nw = TPT >> 5;
  if ((t & 31) == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) red[i * 32 + w] = v[i];
  }
  Group<TPT>::sync(gid);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    for (int k = 0; k < nw; ++k) s += red[i * 32 + k];
    v[i] = s;
  }
  Group<TPT>::sync(gid);
}

// ------------------------------------------------------------------ potentials ---------------
// Evaluates gradient g[d] and dense Hessian H[d][ldh] at qs[d] (all shared memory, group-cooperative) and
// returns this thread's share of the potential energy (sum over the group == V - origin).
// `scr` is a d-vector of shared scratch.  Ends with the results visible to the whole group.
template <int TPT>
__device__ __forceinline__ double pot_eval(const PotDev &P, const double *qs, double *g, double *H, int ldh,
                                           double *scr, double *scr2, int t, int gid, bool fill_h,
                                           bool rotated_h_by_caller = false) {
  const int d = P.d;
  double vpart = 0.0;
  if (P.type == POT_MORSE || P.type == POT_NONHARMONIC) {
    if (fill_h)
      for (int i = t; i < d * d; i += TPT) H[(i / d) * ldh + (i % d)] = 0.0;
    Group<TPT>::sync(gid);
    if (t < d) {
      const double r = qs[t];
      double hd;
      if (P.type == POT_MORSE) {
        if (P.all_harmonic) {
          const double w2 = P.omega[t] * P.omega[t];
          vpart = 0.5 * w2 * r * r;
          g[t] = w2 * r;
          hd = w2;
        } else {
          const double a = P.a[t], D = P.D[t];
          const double e = exp(-a * r);
          vpart = D * (1.0 - e) * (1.0 - e);
          g[t] = 2.0 * a * D * e * (1.0 - e);
          hd = 2.0 * a * a * D * e * (2.0 * e - 1.0);
        }
      } else {
        const double eps = P.eps[t], b = P.b[t];
        const double e1 = exp(-b * r), e2 = exp(-2.0 * b * r);
        vpart = eps / (2.0 * b * b) * (1.0 - e1) * (1.0 - e1) + (1.0 - eps) * 0.5 * r * r;
        g[t] = eps / b * (e1 - e2) + (1.0 - eps) * r;
        hd = eps * (2.0 * e2 - e1) + (1.0 - eps);
      }
      H[t * ldh + t] = hd;
    }
    if (t == 0) vpart -= P.origin;
  } else if (P.type == POT_HARMONIC) {
    if (t < d) scr[t] = qs[t] - P.pos0[t];
    if (fill_h)
      for (int i = t; i < d * d; i += TPT) H[(i / d) * ldh + (i % d)] = P.hess0[i];
    Group<TPT>::sync(gid);
    if (t < d) {
      // H holds hess0 here (filled above, or still there from the first stage): read it from shared memory, not with
      // d uncoalesced global loads per thread
      double hd = 0.0;
      const double *hr = H + t * ldh;
      for (int j = 0; j < d; ++j) hd += hr[j] * scr[j];
      g[t] = P.grad0[t] + hd;
      vpart = scr[t] * P.grad0[t] + 0.5 * scr[t] * hd;
    }
    if (t == 0) vpart += P.e0 - P.origin;
  } else if (P.type == POT_ROTATED_MORSE) {
    // r = Q^T x ; inner Morse ; grad = Q g ; H = Q diag(h) Q^T
    if (t < d) {
      double s = 0.0;
      for (int i = 0; i < d; ++i) s += P.Q[i * d + t] * qs[i];
      double gi, hi;
      if (P.all_harmonic) {
        const double w2 = P.omega[t] * P.omega[t];
        vpart = 0.5 * w2 * s * s;
        gi = w2 * s;
        hi = w2;
      } else {
        const double a = P.a[t], D = P.D[t];
        const double e = exp(-a * s);
        vpart = D * (1.0 - e) * (1.0 - e);
        gi = 2.0 * a * D * e * (1.0 - e);
        hi = 2.0 * a * a * D * e * (2.0 * e - 1.0);
      }
      scr[t] = gi;
      scr2[t] = hi;
    }
    if (t == 0) vpart -= P.origin;
    Group<TPT>::sync(gid);
    if (t < d) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += P.Q[t * d + k] * scr[k];
      g[t] = s;
    }
    // H = Q diag(h) Q^T; the tensor-core kernel forms it itself from scr2 (rotated_hessian_mma, sc_mma.cuh)
    if (!rotated_h_by_caller)
      for (int idx = t; idx < d * d; idx += TPT) {
        const int i = idx / d, j = idx % d;
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += P.Q[i * d + k] * scr2[k] * P.Q[j * d + k];
        H[i * ldh + j] = s;
      }
  }
  Group<TPT>::sync(gid);
  return vpart;
}

// ------------------------------------------------------------------ LU determinant -----------
// Determinant of the dr x dr complex matrix Cm (shared, row-major, ld = dr) by Gaussian elimination with
// implicit partial pivoting (rows are never moved; pivot rows retire from the active list).  Destroys Cm.
// ibuf: 2*dr ints of shared scratch.  Result valid on thread 0 of the group.
template <int TPT>
__device__ __forceinline__ double2 lu_det(double2 *Cm, int dr, int *ibuf, double2 *pivbuf, int t, int gid) {
  int *act = ibuf;        // active (not yet pivoted) physical rows
  int *piv = ibuf + dr;   // piv[k] = physical row chosen at step k
  for (int i = t; i < dr; i += TPT) act[i] = i;
  double2 det = make_double2(1.0, 0.0);
  Group<TPT>::sync(gid);
  for (int k = 0; k < dr; ++k) {
    const int nact = dr - k;
    if (t < 32) {
      // warp 0: pivot = argmax |a[i][k]|^2 over the active rows (ties -> smallest position in act)
      double best = -1.0;
      int bpos = 0;
      for (int pos = t; pos < nact; pos += 32) {
        const double2 v = Cm[act[pos] * dr + k];
        const double m = v.x * v.x + v.y * v.y;
        if (m > best) { best = m; bpos = pos; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
        if (ob > best || (ob == best && op < bpos)) { best = ob; bpos = op; }
      }
      if (t == 0) {
        const int p = act[bpos];
        const double2 pv = Cm[p * dr + k];
        piv[k] = p;
        act[bpos] = act[nact - 1];
        det = cmul(det, pv);
        pivbuf[0] = cinv(pv);
      }
    }
    Group<TPT>::sync(gid);
    if (k + 1 < dr) {
      const int p = piv[k];
      const double2 ip = pivbuf[0];
      const int ncol = dr - k - 1, nrow = nact - 1;
      for (int idx = t; idx < nrow * ncol; idx += TPT) {
        const int i = act[idx / ncol], j = k + 1 + idx % ncol;
        const double2 f = cmul(Cm[i * dr + k], ip);
        const double2 u = Cm[p * dr + j];
        double2 v = Cm[i * dr + j];
        v.x -= f.x * u.x - f.y * u.y;
        v.y -= f.x * u.y + f.y * u.x;
        Cm[i * dr + j] = v;
      }
      Group<TPT>::sync(gid);
    }
  }
  if (t == 0) {
    // parity of the permutation k -> piv[k]
    int swaps = 0;
    for (int k = 0; k < dr; ++k) act[k] = 0;
    for (int k = 0; k < dr; ++k) {
      if (act[k]) continue;
      int len = 0, j = k;
      while (!act[j]) { act[j] = 1; j = piv[j]; ++len; }
      swaps += len - 1;
    }
    if (swaps & 1) { det.x = -det.x; det.y = -det.y; }
  }
  return det;
}

// CTA-wide variant (TPT a multiple of 32, dr <= 64): one barrier per column.
//   rows -> warps (row i belongs to warp i % NW), columns -> lanes: no integer division, conflict-free row
//   accesses; the pivot of column k+1 is found while column k is eliminated (each warp reduces its rows with
//   redux.sync on a packed (magnitude, row) key, the NW partial keys are combined by every warp after the
//   barrier); retired rows are tracked in a 64-bit register mask, the permutation parity with popcounts.
// wkey: 2 x 32 unsigned of shared scratch (double-buffered partial keys). Result valid on every thread.
__device__ __forceinline__ unsigned pivot_key(double2 v, int row) {
  const double m = v.x * v.x + v.y * v.y;
  return (static_cast<unsigned>(__double2hiint(m)) & ~63u) | static_cast<unsigned>(row);
}

template <int TPT>
__device__ __forceinline__ double2 lu_det_cta(double2 *Cm, int dr, unsigned *wkey, int t) {
  constexpr int NW = TPT / 32;
  constexpr int MAXR = (64 + NW - 1) / NW;   // rows per warp
  const int lane = t & 31, warp = t >> 5;
  unsigned long long done = 0ull;
  double2 det = make_double2(1.0, 0.0);
  int inversions = 0;
  {
    // partial keys of column 0
    unsigned key = 0u;
    const int i = warp + NW * lane;
    if (lane < MAXR && i < dr) key = pivot_key(Cm[i * dr], i);
    key = __reduce_max_sync(0xffffffffu, key);
    if (lane == 0) wkey[warp] = key;
  }
  __syncthreads();
  for (int k = 0; k < dr; ++k) {
    // loads that do not depend on the pivot are issued first: partial keys, this warp's rows
    const unsigned *wk = wkey + (k & 1) * 32;
    unsigned kk = (lane < NW) ? wk[lane] : 0u;
    const int j0 = k + 1 + lane, j1 = j0 + 32;
    const bool c0 = j0 < dr, c1 = j1 < dr;
    double2 ak[MAXR], v0[MAXR], v1[MAXR];
#pragma unroll
    for (int m = 0; m < MAXR; ++m) {
      const int i = warp + NW * m;
      ak[m] = v0[m] = v1[m] = make_double2(0.0, 0.0);
      if (i < dr) {
        const double2 *row = Cm + i * dr;
        ak[m] = row[k];
        if (c0) v0[m] = row[j0];
        if (c1) v1[m] = row[j1];
      }
    }
    kk = __reduce_max_sync(0xffffffffu, kk);
    const int p = static_cast<int>(kk & 63u);
    inversions += __popcll(done >> p);          // earlier pivots with a larger row index
    done |= 1ull << p;
    const double2 *prow = Cm + p * dr;
    const double2 pv = prow[k];
    double2 u0 = make_double2(0.0, 0.0), u1 = u0;
    if (c0) u0 = prow[j0];
    if (c1) u1 = prow[j1];
    det = cmul(det, pv);
    if (k + 1 == dr) break;
    const double rn = 1.0 / (pv.x * pv.x + pv.y * pv.y);
    const double2 ip = make_double2(pv.x * rn, -pv.y * rn);
    unsigned nkey = 0u;
#pragma unroll
    for (int m = 0; m < MAXR; ++m) {
      const int i = warp + NW * m;
      if (i < dr && !((done >> i) & 1ull)) {
        const double2 f = cmul(ak[m], ip);
        double2 *row = Cm + i * dr;
        if (c0) {
          double2 v = v0[m];
          v.x -= f.x * u0.x - f.y * u0.y;
          v.y -= f.x * u0.y + f.y * u0.x;
          row[j0] = v;
          if (lane == 0) nkey = max(nkey, pivot_key(v, i));
        }
        if (c1) {
          double2 v = v1[m];
          v.x -= f.x * u1.x - f.y * u1.y;
          v.y -= f.x * u1.y + f.y * u1.x;
          row[j1] = v;
        }
      }
    }
    if (lane == 0) wkey[((k + 1) & 1) * 32 + warp] = nkey;
    __syncthreads();
  }
  if (inversions & 1) { det.x = -det.x; det.y = -det.y; }
  return det;
}

// named barrier of the blocked LU kernels (sc_lu.cuh): id 0 is __syncthreads
__device__ __forceinline__ unsigned lu_key(double m, int row) {
  return (static_cast<unsigned>(__double2hiint(m)) & ~63u) + 64u + static_cast<unsigned>(row);
}

template <int BAR_ID, int NTHR>
__device__ __forceinline__ void lu_bar() {
  if (BAR_ID == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"n"(BAR_ID), "n"(NTHR) : "memory");
}

// ------------------------------------------------------------------ prefactor assembly -------
// Cm (dr x dr complex) = 1/2 [ L1 Mqq R1 + L2 Mpp R2 - i L1 Mqp R2 + i L2 Mpq R1 ]     (propagators.py:969-994)
// Ub = [Mqq|Mqp], Vb = [Mpq|Mpp] (d x 2d, ld = ldu) in shared memory; T: shared scratch d x dr.
template <int TPT>
__device__ __forceinline__ void prefactor_assemble(const EngDev &E, const double *Ub, const double *Vb, int ldu,
                                                   double2 *Cm, int ldc, double *T, int t, int gid) {
  const int d = E.d, dr = E.dr;
  if (E.diag) {
    for (int idx = t; idx < d * d; idx += TPT) {
      const int a = idx / d, b = idx % d;
      const double mqq = Ub[a * ldu + b], mqp = Ub[a * ldu + d + b], mpq = Vb[a * ldu + b], mpp = Vb[a * ldu + d + b];
      const double sa = E.sgt[a], isa = E.isgt[a], sb = E.sgi[b], isb = E.isgi[b];
      Cm[a * ldc + b] = make_double2(0.5 * (sa * mqq * isb + isa * mpp * sb), 0.5 * (-sa * mqp * sb + isa * mpq * isb));
    }
    Group<TPT>::sync(gid);
    return;
  }
  for (int blk = 0; blk < 4; ++blk) {
    // blk 0: Mqq (L1,R1) re ; 1: Mpp (L2,R2) re ; 2: Mqp (L1,R2) -im ; 3: Mpq (L2,R1) +im
    const double *Mb = (blk == 0 || blk == 2) ? Ub : Vb;
    const int off = (blk == 1 || blk == 2) ? d : 0;
    const double *L = (blk == 0 || blk == 2) ? E.L1 : E.L2;
    const double *R = (blk == 0 || blk == 3) ? E.R1 : E.R2;
    for (int idx = t; idx < d * dr; idx += TPT) {
      const int a = idx / dr, bp = idx % dr;
      double s = 0.0;
      for (int b = 0; b < d; ++b) s += Mb[a * ldu + off + b] * __ldg(R + b * dr + bp);
      T[idx] = s;
    }
    Group<TPT>::sync(gid);
    for (int idx = t; idx < dr * dr; idx += TPT) {
      const int ap = idx / dr, bp = idx % dr;
      double s = 0.0;
      for (int a = 0; a < d; ++a) s += __ldg(L + ap * d + a) * T[a * dr + bp];
      s *= 0.5;
      double2 v = (blk == 0) ? make_double2(0.0, 0.0) : Cm[ap * ldc + bp];
      if (blk < 2) v.x += s;
      else if (blk == 2) v.y -= s;
      else v.y += s;
      Cm[ap * ldc + bp] = v;
    }
    Group<TPT>::sync(gid);
  }
}

// sqrt branch tracking (propagators.py:1045-1047)
__device__ __forceinline__ double track_sign(double sign, double2 zprev, double2 z) {
  return (zprev.x < 0.0 && z.x < 0.0 && zprev.y * z.y < 0.0) ? -sign : sign;
}

// ------------------------------------------------------------------ correlation terms --------
// Per-thread partial sums (thread a < d handles component a) of everything the HK contributions need:
//  v[0] = -1/2 dq A dq - 1/2 dp B dp        (real part of the overlap exponent, bra = current point)
//  v[1] = -p0.dq + dq C dp                  (imaginary part)
//  v[2] = (q0-Q).wR      v[3] = (P-p0).wG   (current-time NAC factor)
//  v[4] = (q0-qi).wR     v[5] = (pi-p0).wG  (initial NAC factor)
// q, p: shared (current); zt: global (initial point of this trajectory); dqv, dpv: shared scratch (d)
template <int TPT>
__device__ __forceinline__ void corr_terms(const EngDev &E, const double *q, const double *p, const double *zt,
                                           double *dqv, double *dpv, double (&v)[6], int t, int gid) {
  const int d = E.d;
  if (t < d) { dqv[t] = E.q0[t] - q[t]; dpv[t] = E.p0[t] - p[t]; }
  Group<TPT>::sync(gid);
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = 0.0;
  if (t < d) {
    const double dq = dqv[t], dp = dpv[t];
    if (E.diag) {
      v[0] = -0.5 * (dq * E.otA[t] * dq + dp * E.otB[t] * dp);
      v[1] = -E.p0[t] * dq + dq * E.otC[t] * dp;
    } else {
      // column t of A, B, C (coalesced across the threads): sum_t x_t (A^T x)_t = x A x, and
      // sum_t dp_t (C^T dq)_t = dq C dp
      double sa = 0.0, sb = 0.0, sc_ = 0.0;
      for (int j = 0; j < d; ++j) {
        sa += __ldg(E.otA + j * d + t) * dqv[j];
        sb += __ldg(E.otB + j * d + t) * dpv[j];
        sc_ += __ldg(E.otC + j * d + t) * dqv[j];
      }
      v[0] = -0.5 * (dq * sa + dp * sb);
      v[1] = -E.p0[t] * dq + dp * sc_;
    }
    const double wr = E.wR[t], wg = E.wG[t];
    v[2] = dq * wr;
    v[3] = -dp * wg;
    v[4] = (E.q0[t] - zt[t]) * wr;
    v[5] = (zt[d + t] - E.p0[t]) * wg;
  }
}

// thread-0 epilogue: complex contributions of one trajectory to C_auto and k_ic (weights folded into wvi)
__device__ __forceinline__ void corr_finish(const EngDev &E, const double (&v)[6], double S, double2 c, double sign,
                                            double2 wvi, double2 &cauto, double2 &kic) {
  // vt = fac exp(v0 + i v1); contribution = conj(vt) * wvi * (sign c) * exp(i S)
  const double2 e = cexp(v[0], S - v[1]);
  double2 cq = cmul(make_double2(E.ot_fac * e.x, E.ot_fac * e.y), wvi);
  cq = cmul(cq, make_double2(sign * c.x, sign * c.y));
  cauto = cq;
  const double2 nacQ = make_double2(v[2], -(E.p0n1 + v[3]));
  const double2 nacq = make_double2(v[4], (E.p0n1 + v[5]));
  kic = cmul(cmul(nacQ, nacq), cq);
}

}  // namespace sc
