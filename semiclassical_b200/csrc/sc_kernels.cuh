// sc_kernels.cuh -- fused Herman-Kluk step kernels.
//
// k_hk_generic<TPT, EPT>: any d <= 64, any potential, dense or diagonal Gamma.  One group (warp or CTA) per
// trajectory, state resident in shared memory for K fused time steps; the H * [Mqq|Mqp] products are plain
// DFMA dot products.  This is the correctness baseline and the small-d production path.
//
// RK4 bookkeeping (classical RK4 of propagators.py:114-119 written for the block structure of the equations
// of motion, propagators.py:342-357).  With U = [Mqq|Mqp], V = [Mpq|Mpp], kv_s = -H(q_s) U_s :
//   U_2 = U + h/2 V/m                 U_3 = U + h/2 (V + h/2 kv_1)/m      U_4 = U + h (V + h/2 kv_2)/m
//   U'  = U + h V/m + h^2/6 (kv_1 + kv_2 + kv_3)/m
//   V'  = V + h/6 (kv_1 + 2 kv_2 + 2 kv_3 + kv_4)
// which needs only two register accumulators per element (R1 = running sum for U', R2 = V' - U' part) next
// to the product in flight; kv_{s-1} is recovered from them when U_{s+1} is formed.
#pragma once
#include "sc_device.cuh"
#ifndef SC_MAX_DIM
#define SC_MAX_DIM 96
#endif

namespace sc {

enum KernelMode { MODE_STEP = 0, MODE_INIT = 1, MODE_CORR = 2, MODE_TRACK = 3 };  // TRACK: prefactor + branch tracking only

struct SmemLayout {
  int ldu, ldh, dpad;
  int off_Ub, off_Vb, off_Us, off_H, off_vec, off_red, off_int, off_lu, off_acc, total;  // in doubles, per group
};

__host__ __device__ inline SmemLayout make_layout(int d, int dr, int ldu, int ldh, int mma_kmax = 0) {
  SmemLayout L;
  L.ldu = ldu;
  L.ldh = ldh;
  L.dpad = (d + 1) & ~1;
  int o = 0;
  L.off_Ub = o; o += d * ldu;
  L.off_Vb = o; o += d * ldu;
  int us = ((d + 3) & ~3) * ldu;           // rows padded to the k-step of the tensor-core variant
  if (us < 2 * dr * (dr + 1)) us = 2 * dr * (dr + 1);   // Us doubles as the complex prefactor matrix (odd row stride)
  L.off_Us = o; o += us;
  int hs = ((d + 7) & ~7) * ldh;            // rows padded to 8 for the tensor-core variant
  if (hs < d * dr) hs = d * dr;             // H doubles as the T scratch of the prefactor assembly
  L.off_H = o; o += hs;
  L.off_vec = o; o += 14 * L.dpad + 8;      // q, p, qs, g, scr, scr2, dqv, dpv, 4 Hessian diagonals, 2 stashes
  L.off_red = o; o += 8 * 32;               // cross-warp reduction scratch
  L.off_int = o; o += (dr < 32 ? 38 : (dr + 6) & ~1);  // LU bookkeeping (2 dr ints | 64 keys) + pivot inverse (double2)
  o = (o + 1) & ~1;
  L.off_lu = o; L.off_acc = o;
  (void)mma_kmax;
  L.total = (o + 1) & ~1;
  return L;
}

template <int TPT, int EPT>
__global__ void __launch_bounds__(TPT == 32 ? 128 : TPT)
k_hk_generic(EngDev E, PotDev P, double h, int nsteps, int mode, double *partials, SmemLayout L) {
  extern __shared__ __align__(16) double smem[];
  constexpr int G = (TPT == 32) ? 4 : 1;
  const int gid = (TPT == 32) ? (threadIdx.x >> 5) : 0;
  const int t = (TPT == 32) ? (threadIdx.x & 31) : threadIdx.x;
  const int gg = blockIdx.x * G + gid, NG = gridDim.x * G;
  double *base = smem + (size_t)gid * L.total;
  double *Ub = base + L.off_Ub, *Vb = base + L.off_Vb, *Us = base + L.off_Us, *H = base + L.off_H;
  double *vec = base + L.off_vec, *red = base + L.off_red;
  int *ibuf = reinterpret_cast<int *>(base + L.off_int);
  const int d = E.d, dr = E.dr, ldu = L.ldu, ldh = L.ldh, dp = L.dpad, W = 2 * d, NE = 2 * d * d;
  double *q = vec, *p = vec + dp, *qs = vec + 2 * dp, *g = vec + 3 * dp, *scr = vec + 4 * dp, *scr2 = vec + 5 * dp;
  double *dqv = vec + 6 * dp, *dpv = vec + 7 * dp;
  double2 *pivbuf = reinterpret_cast<double2 *>(ibuf + ((2 * dr + 3) & ~3));
  double2 *Cm = reinterpret_cast<double2 *>(Us);
  const double im_t = (t < d) ? P.imass[t] : 0.0;

  for (int traj = gg; traj < E.n; traj += NG) {
    double *rec = E.rec + (size_t)traj * E.rs;
    // ---- load the record: [q p S pad | U | V]
    if (t < d) { q[t] = rec[t]; p[t] = rec[d + t]; }
    double S = rec[2 * d];
    for (int idx = t; idx < NE; idx += TPT) {
      const int a = idx / W, b = idx % W;
      Ub[a * ldu + b] = rec[E.qps + idx];
      Vb[a * ldu + b] = rec[E.qps + NE + idx];
    }
    double2 c2 = E.c2[traj], cc = E.c[traj];
    double sign = E.sign[traj];
    Group<TPT>::sync(gid);

    const int nloop = (mode == MODE_STEP) ? nsteps : 1;
    for (int step = 0; step < nloop; ++step) {
      double e4 = 0.0, accS = 0.0;
      if (mode == MODE_STEP) {
        // ================= one classical RK4 step =================
        double R1[EPT], R2[EPT];
        double qa = 0, pa = 0, qsa = 0, psa = 0, accq = 0, accp = 0;
        if (t < d) { qa = q[t]; pa = p[t]; qsa = qa; psa = pa; qs[t] = qa; }
        Group<TPT>::sync(gid);
        double vpart = pot_eval<TPT>(P, qs, g, H, ldh, scr, scr2, t, gid, true);
#pragma unroll 1
        for (int s = 1; s <= 4; ++s) {
          const double cnext = (s == 3) ? h : 0.5 * h;  // c_{s+1}
          const double wgt = (s == 1 || s == 4) ? 1.0 : 2.0;
          // ---- phase A: kv = -H U_s
          const double *Ucur = (s == 1) ? Ub : Us;
          double kv[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const int idx = t + e * TPT;
            double acc = 0.0;
            if (idx < NE) {
              const int a = idx / W, b = idx % W;
              const double *hr = H + a * ldh;
              const double *uc = Ucur + b;
              for (int k = 0; k < d; ++k) acc = fma(hr[k], uc[k * ldu], acc);
            }
            kv[e] = -acc;
          }
          double kq = 0, kp = 0;
          if (t < d) {
            kq = psa * im_t;
            kp = -g[t];
            const double tk = 0.5 * psa * psa * im_t;
            accS += wgt * (tk - vpart);
            if (s == 4) e4 = tk + vpart;
          }
          Group<TPT>::sync(gid);
          // ---- phase B: accumulators, next stage operand
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const int idx = t + e * TPT;
            if (idx < NE) {
              const int a = idx / W, b = idx % W;
              const double ima = P.imass[a];
              const double ub = Ub[a * ldu + b], vb = Vb[a * ldu + b];
              if (s == 1) {
                R1[e] = kv[e];
                R2[e] = 0.0;
                Us[a * ldu + b] = ub + 0.5 * h * vb * ima;                       // U_2
              } else if (s == 2) {
                Us[a * ldu + b] = ub + 0.5 * h * (vb + 0.5 * h * R1[e]) * ima;   // U_3 (R1 == kv_1)
                R1[e] += kv[e];
                R2[e] = kv[e];
              } else if (s == 3) {
                Us[a * ldu + b] = ub + h * (vb + 0.5 * h * R2[e]) * ima;         // U_4 (R2 == kv_2)
                R1[e] += kv[e];
                R2[e] += kv[e];
              } else {
                R2[e] += kv[e];
                Ub[a * ldu + b] = ub + h * vb * ima + (h * h / 6.0) * R1[e] * ima;
                Vb[a * ldu + b] = vb + (h / 6.0) * (R1[e] + R2[e]);
              }
            }
          }
          if (t < d) {
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
          Group<TPT>::sync(gid);
          if (s < 4) vpart = pot_eval<TPT>(P, qs, g, H, ldh, scr, scr2, t, gid, P.type == POT_ROTATED_MORSE);
        }
      }
      // ================= prefactor + branch tracking =================
      double2 det = make_double2(0.0, 0.0);
      if (mode != MODE_CORR) {
        prefactor_assemble<TPT>(E, Ub, Vb, ldu, Cm, dr, H, t, gid);
        if (TPT == 32) det = lu_det<TPT>(Cm, dr, ibuf, pivbuf, t, gid);
        else det = lu_det_cta<(TPT == 32 ? 64 : TPT)>(Cm, dr, reinterpret_cast<unsigned *>(ibuf), t);
      }
      // ================= correlation contributions =================
      double v8[8];
      {
        double v6[6];
        corr_terms<TPT>(E, q, p, E.zt + (size_t)traj * 2 * d, dqv, dpv, v6, t, gid);
#pragma unroll
        for (int i = 0; i < 6; ++i) v8[i] = v6[i];
        v8[6] = accS;
        v8[7] = e4;
      }
      group_reduce<TPT, 8>(v8, red, t, gid);
      if (t == 0) {
        if (mode == MODE_STEP || mode == MODE_TRACK) {
          S += h / 6.0 * v8[6];
          sign = track_sign(sign, c2, det);
          c2 = det;
          cc = csqrt_principal(det);
        } else if (mode == MODE_INIT) {
          sign = 1.0;
          c2 = det;
          cc = csqrt_principal(det);
        }
        if (mode != MODE_INIT && mode != MODE_TRACK) {
          double2 ca, ki;
          const double v6[6] = {v8[0], v8[1], v8[2], v8[3], v8[4], v8[5]};
          corr_finish(E, v6, S, cc, sign, E.wvi[traj], ca, ki);
          double *row = partials + ((size_t)gg * nloop + step) * 5;
          row[0] += ca.x; row[1] += ca.y; row[2] += ki.x; row[3] += ki.y; row[4] += v8[7];
        }
      }
      if (E.snap != nullptr && mode == MODE_STEP) {
        // snapshot of the new time for the fused Walton-Manolopoulos launch (sc_wm.cuh: k_wm_fused)
        const size_t item = (size_t)step * E.n + traj;
        double *sr = E.snap + item * E.rs;
        if (t < d) { sr[t] = q[t]; sr[d + t] = p[t]; }
        for (int idx = t; idx < NE; idx += TPT) {
          const int a = idx / W, b = idx % W;
          sr[E.qps + idx] = Ub[a * ldu + b];
          sr[E.qps + NE + idx] = Vb[a * ldu + b];
        }
        if (t == 0) {
          sr[2 * d] = S;
          E.snap_c[item] = cc;
          E.snap_sign[item] = sign;
        }
      }
      Group<TPT>::sync(gid);
    }
    // ---- write back
    if (mode == MODE_STEP) {
      if (t < d) { rec[t] = q[t]; rec[d + t] = p[t]; }
      if (t == 0) rec[2 * d] = S;
      for (int idx = t; idx < NE; idx += TPT) {
        const int a = idx / W, b = idx % W;
        rec[E.qps + idx] = Ub[a * ldu + b];
        rec[E.qps + NE + idx] = Vb[a * ldu + b];
      }
    }
    if (t == 0 && mode != MODE_CORR) {
      E.c2[traj] = c2;
      E.c[traj] = cc;
      E.sign[traj] = sign;
    }
    Group<TPT>::sync(gid);
  }
}

// deterministic second pass: sums the per-group partial rows in a fixed order.
// partials: (ngroups, nsteps, 5) -> out (nsteps, 5); correlation sums scaled by inv_norm, energy by 1/n
__global__ void k_reduce_partials(const double *partials, int ngroups, int nsteps, double inv_norm, double inv_n,
                                  double *out) {
  const int k = blockIdx.x, j = threadIdx.x >> 5, lane = threadIdx.x & 31;  // 5 warps, one per column
  if (j >= 5) return;
  double s = 0.0;
  for (int gidx = lane; gidx < ngroups; gidx += 32) s += partials[((size_t)gidx * nsteps + k) * 5 + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[k * 5 + j] = s * (j < 4 ? inv_norm : inv_n);
}

// ensemble installation: zi (2d, n) batch-last -> trajectory-major records with Mqq = Mpp = 1, S = 0;
// also the time-independent overlap <qi,pi,Gi|q0,p0,G0> / (probi (2 pi)^d)   (propagators.py:581-603, 795, 837)
// One CTA per trajectory (grid-stride): record writes are coalesced, the overlap is reduced by warp 0.
__global__ void k_init_records(EngDev E, const double *zi, const double *probi, const double *oiA, const double *oiB,
                               const double *oiC, double oi_fac, double inv2pid, double *zt_out, double2 *wvi_out) {
  __shared__ double dq[SC_MAX_DIM], dpv[SC_MAX_DIM];
  const int d = E.d, n = E.n, NE = 2 * d * d, W = 2 * d, t = threadIdx.x;
  for (int traj = blockIdx.x; traj < n; traj += gridDim.x) {
    double *rec = E.rec + (size_t)traj * E.rs;
    double *zt = zt_out + (size_t)traj * 2 * d;
    for (int k = t; k < 2 * d; k += blockDim.x) {
      const double z = zi[(size_t)k * n + traj];
      zt[k] = z;
      rec[k] = z;
      if (k < d) dq[k] = E.q0[k] - z;
      else dpv[k - d] = E.p0[k - d] - z;
    }
    if (t == 0) rec[2 * d] = 0.0;
    for (int idx = t; idx < NE; idx += blockDim.x) {
      const int a = idx / W, b = idx % W;
      rec[E.qps + idx] = (b == a) ? 1.0 : 0.0;            // U = [1 | 0]
      rec[E.qps + NE + idx] = (b == d + a) ? 1.0 : 0.0;   // V = [0 | 1]
    }
    __syncthreads();
    if (t < 32) {
      // overlap with bra = (qi, pi, Gamma_i), ket = (q0, p0, Gamma_0)
      double re = 0.0, im = 0.0;
      for (int a = t; a < d; a += 32) {
        double sa = 0.0, sb = 0.0, sc_ = 0.0;
        if (E.diag) {
          sa = oiA[a] * dq[a]; sb = oiB[a] * dpv[a]; sc_ = oiC[a] * dpv[a];
        } else {
          for (int j = 0; j < d; ++j) {
            sa += oiA[a * d + j] * dq[j]; sb += oiB[a * d + j] * dpv[j]; sc_ += oiC[a * d + j] * dpv[j];
          }
        }
        re += -0.5 * (dq[a] * sa + dpv[a] * sb);
        im += -E.p0[a] * dq[a] + dq[a] * sc_;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
      }
      if (t == 0) {
        const double2 e = cexp(re, im);
        const double w = oi_fac * inv2pid / probi[traj];
        wvi_out[traj] = make_double2(w * e.x, w * e.y);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ ensemble sampling on the device (propagators.py:533-555)
// Philox4x32-10 counter-based generator (Salmon et al., SC'11): key = seed, counter = (trajectory index, draw index); two
// 53-bit uniforms per call -> two standard normals by Box-Muller.  Counter-based, so the ensemble is a pure function of
// (seed, global trajectory index): any sharding of the index range over ranks draws the same global ensemble.
__device__ __forceinline__ void philox4x32_10(unsigned (&c)[4], unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// zi (2d, n) batch-last, probi (n):  x ~ N(0,1)^(2 d'),  z = z0 + (Lz^-1)^T x,  P = detLz / (2 pi)^d exp(-x.x / 2)
// iLz = blockdiag(iLq, iLp): iLq, iLp (d' x d) row-major.  One warp per trajectory: lane l draws the normals l, l + 32, ...
// and owns the components l, l + 32, ... of q and p.
__global__ void __launch_bounds__(128)
k_sample_ensemble(int d, int dr, int n, long long index0, unsigned long long seed, const double *__restrict__ iLq,
                  const double *__restrict__ iLp, const double *__restrict__ q0, const double *__restrict__ p0, double pfac,
                  double *__restrict__ zi, double *__restrict__ probi) {
  constexpr int NE = (SC_MAX_DIM + 31) / 32;                    // components / normals per lane and half
  const int lane = threadIdx.x & 31;
  const int traj = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (traj >= n) return;
  const unsigned long long gi = (unsigned long long)(index0 + traj);
  double xq[NE], xp[NE], x2 = 0.0;
#pragma unroll
  for (int k = 0; k < NE; ++k) {
    xq[k] = xp[k] = 0.0;
    const int j = lane + 32 * k;
    if (j < dr) {
      unsigned c[4] = {(unsigned)gi, (unsigned)(gi >> 32), (unsigned)j, 0x5eedu};
      philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
      // two uniforms in (0, 1]: 53 random bits each
      const double u1 = ((double)(((unsigned long long)c[0] << 21) ^ (c[1] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
      const double u2 = ((double)(((unsigned long long)c[2] << 21) ^ (c[3] >> 11)) + 1.0) * (1.0 / 9007199254740992.0);
      const double rad = sqrt(-2.0 * log(u1));
      double sn, cs;
      sincospi(2.0 * u2, &sn, &cs);
      xq[k] = rad * cs;                                          // normal j of the position block
      xp[k] = rad * sn;                                          // normal j of the momentum block
      x2 += xq[k] * xq[k] + xp[k] * xp[k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x2 += __shfl_xor_sync(0xffffffffu, x2, o);
  double zq[NE], zp[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) {
    const int a = lane + 32 * k;
    zq[k] = a < d ? q0[a] : 0.0;
    zp[k] = a < d ? p0[a] : 0.0;
  }
  for (int j = 0; j < dr; ++j) {
    const int src = j & 31, kk = j >> 5;
    double vq = xq[0], vp = xp[0];
#pragma unroll
    for (int k = 1; k < NE; ++k)
      if (kk == k) { vq = xq[k]; vp = xp[k]; }
    vq = __shfl_sync(0xffffffffu, vq, src);
    vp = __shfl_sync(0xffffffffu, vp, src);
#pragma unroll
    for (int k = 0; k < NE; ++k) {
      const int a = lane + 32 * k;
      if (a < d) {
        zq[k] = fma(iLq[(size_t)j * d + a], vq, zq[k]);
        zp[k] = fma(iLp[(size_t)j * d + a], vp, zp[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NE; ++k) {
    const int a = lane + 32 * k;
    if (a < d) {
      zi[(size_t)a * n + traj] = zq[k];
      zi[(size_t)(d + a) * n + traj] = zp[k];
    }
  }
  if (lane == 0) probi[traj] = pfac * exp(-0.5 * x2);
}

// prefactor of trajectory 0 -> all trajectories (initial conditions: identical monodromy matrices)
__global__ void k_broadcast_prefactor(double2 *c2, double2 *c, double *sign, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > 0 && i < n) {
    c2[i] = c2[0];
    c[i] = c[0];
    sign[i] = sign[0];
  }
}

// layout conversion: records <-> the reference's y (2d+4d^2+1, n)
__global__ void k_export_state(EngDev E, double *y, int to_y) {
  const int d = E.d, n = E.n, d2 = d * d, L = 2 * d + 4 * d2 + 1;
  const size_t total = (size_t)L * n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / n), traj = (int)(i % n);
    double *rec = E.rec + (size_t)traj * E.rs;
    int off;
    if (row < 2 * d) off = row;
    else if (row == L - 1) off = 2 * d;
    else {
      const int r = row - 2 * d, blk = r / d2, a = (r % d2) / d, b = r % d;  // blk: Mqq, Mqp, Mpq, Mpp
      off = E.qps + ((blk >= 2) ? 2 * d2 : 0) + a * 2 * d + ((blk & 1) ? d : 0) + b;
    }
    if (to_y) y[i] = rec[off];
    else rec[off] = y[i];
  }
}

}  // namespace sc
