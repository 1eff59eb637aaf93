// sc_chunk.cuh -- column-chunked Herman-Kluk step for large d with diagonal width matrices (the headline path).
//
// The 2d columns of the monodromy blocks are independent linear ODEs driven by the same Hessians
// (propagators.py:342-357: d/dt [Mqq;Mpq][:,b] and d/dt [Mqp;Mpp][:,b] only involve column b), and with diagonal
// Gamma column b of the prefactor matrix (propagators.py:969-986) needs exactly column b of the four blocks.  The
// step is therefore split into three throughput kernels:
//
//   k_rk4_chunk   work item = (trajectory, chunk of nb columns b): a 128-thread CTA keeps U = [Mqq|Mqp][:, chunk],
//                 V = [Mpq|Mpp][:, chunk] (d x 2nb each) and the dense Hessian in ~72 KB of shared memory for K time
//                 steps -> THREE independent CTAs per SM, so one CTA's elementwise / barrier / potential phases
//                 overlap the other CTAs' DMMA.  H U_s runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64), the RK4
//                 accumulators stay in registers, and the stage operand U_s is formed IN PLACE:
//                     U_2 = U + h/2 V/m,  U_3 = U_2 + h^2/4 kv_1/m,  U_4 = U_3 + [h/2 V + h^2/2 kv_2 - h^2/4 kv_1]/m,
//                     U'  = U_4 + [h^2/6 (kv_1+kv_2+kv_3) - h^2/2 kv_2]/m,   V' = V + h/6 (kv_1+2kv_2+2kv_3+kv_4)
//                 (kv_s = -H_s U_s; algebraically the classical RK4 of propagators.py:114-119).  Every step the CTA
//                 writes its nb columns of the complex prefactor matrix to a global scratch; chunk 0 also writes the
//                 overlap / NAC partial sums, the action and T+V.
//   k_lu_batch    (sc_lu_batch.cuh) determinants of all (step, trajectory) matrices of the batch.
//   k_hk_finish   per trajectory, in time order: sqrt branch tracking (propagators.py:1045-1047), contributions to
//                 C_auto and k_ic (propagators.py:784-911), deterministic per-block partial sums.
//
// Every chunk CTA integrates (q, p) itself (separable potentials: d threads, no dependence on M); that is 3x
// redundant work of O(d) per step against O(d^3)/3 of monodromy work.
#pragma once
#include "sc_mma.cuh"

namespace sc {

struct ChunkLayout {
  int nc, nb, ldu, ldh, dk;                 // chunks per trajectory, columns b per chunk, leading dimensions, K padding
  int off_U, off_V, off_H, off_vec, total;  // doubles
};

__host__ __device__ inline ChunkLayout make_chunk_layout(int d) {
  ChunkLayout L;
  L.nc = (d <= 60) ? 3 : 4;
  L.nb = (d + L.nc - 1) / L.nc;
  int w = 2 * L.nb;
  w = (w + 7) & ~7;                          // whole n-tiles
  L.ldu = w;
  while (L.ldu % 16 != 8) L.ldu += 8;
  L.dk = (d + 3) & ~3;
  L.ldh = L.dk;
  while (L.ldh % 16 != 4 && L.ldh % 16 != 12) L.ldh += 4;
  int o = 0;
  L.off_U = o; o += L.dk * L.ldu;
  L.off_V = o; o += d * L.ldu;
  L.off_H = o; o += d * L.ldh;
  L.off_vec = o; o += 12 * ((d + 1) & ~1);   // Hessian diagonals of 4 stages (double buffered), sgt/2, isgt/2, sgi, isgi
  L.total = (o + 1) & ~1;
  return L;
}

constexpr int CHUNK_THREADS = 128;
constexpr int CHUNK_WN = 5;   // n-tiles per warp: 2 nb <= 40

// (q, p, S) of the separable model for K time steps, one warp per trajectory (lane owns the modes lane, lane+32):
// classical RK4 on (q, p) (propagators.py:114-119, 361-368), the four Hessian diagonals of every step for the
// matrix kernel, and per step the overlap / NAC partial sums, the action and T+V of the 4th stage point.
//   hd : (nsteps, ntb, 4, dp) Hessian diagonals      aux : (nsteps, ntb, 8)
__global__ void __launch_bounds__(128)
k_qp_path(EngDev E, PotDev P, double h, int nsteps, int traj0, int ntb, double *__restrict__ hd, double *__restrict__ aux) {
  const int lane = threadIdx.x & 31;
  const int tl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tl >= ntb) return;
  const int traj = traj0 + tl, d = E.d, dp = (d + 1) & ~1;
  double *rec = E.rec + (size_t)traj * E.rs;
  const double *zt = E.zt + (size_t)traj * 2 * d;
  double q[2], p[2], im[2], q0[2], p0[2], oA[2], oB[2], oC[2], wr[2], wg[2], ci4[2], ci5[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int a = lane + 32 * k;
    const bool ok = a < d;
    q[k] = ok ? rec[a] : 0.0;
    p[k] = ok ? rec[d + a] : 0.0;
    im[k] = ok ? P.imass[a] : 0.0;
    q0[k] = ok ? E.q0[a] : 0.0;
    p0[k] = ok ? E.p0[a] : 0.0;
    oA[k] = ok ? E.otA[a] : 0.0;
    oB[k] = ok ? E.otB[a] : 0.0;
    oC[k] = ok ? E.otC[a] : 0.0;
    wr[k] = ok ? E.wR[a] : 0.0;
    wg[k] = ok ? E.wG[a] : 0.0;
    ci4[k] = ok ? (q0[k] - zt[a]) * wr[k] : 0.0;        // initial-point NAC factors (time independent)
    ci5[k] = ok ? (zt[d + a] - p0[k]) * wg[k] : 0.0;
  }
  double S = rec[2 * d];
  for (int step = 0; step < nsteps; ++step) {
    double *hrow = hd + ((size_t)step * ntb + tl) * 4 * dp;
    double v8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int a = lane + 32 * k;
      if (a < d) {
        const double qa = q[k], pa = p[k];
        double qsa = qa, psa = pa, accq = 0.0, accp = 0.0, accS = 0.0, e4 = 0.0;
#pragma unroll
        for (int s = 1; s <= 4; ++s) {
          const double cnext = (s == 3) ? h : 0.5 * h;
          const double wgt = (s == 1 || s == 4) ? 1.0 : 2.0;
          double gt, hdg;
          const double vpart = pot_local_vals(P, a, qsa, gt, hdg);
          hrow[(s - 1) * dp + a] = hdg;
          const double kq = psa * im[k], kp = -gt;
          const double tk = 0.5 * psa * psa * im[k];
          accS += wgt * (tk - vpart);
          if (s == 4) e4 = tk + vpart;
          accq += wgt * kq;
          accp += wgt * kp;
          if (s < 4) {
            qsa = qa + cnext * kq;
            psa = pa + cnext * kp;
          }
        }
        q[k] = qa + h / 6.0 * accq;
        p[k] = pa + h / 6.0 * accp;
        const double dq = q0[k] - q[k], dpp = p0[k] - p[k];
        v8[0] += -0.5 * (dq * oA[k] * dq + dpp * oB[k] * dpp);
        v8[1] += -p0[k] * dq + dq * oC[k] * dpp;
        v8[2] += dq * wr[k];
        v8[3] += -dpp * wg[k];
        v8[4] += ci4[k];
        v8[5] += ci5[k];
        v8[6] += accS;
        v8[7] += e4;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v8[i] += __shfl_xor_sync(0xffffffffu, v8[i], o);
    }
    S += h / 6.0 * v8[6];
    if (lane == 0) {
      double *ax = aux + ((size_t)step * ntb + tl) * 8;
#pragma unroll
      for (int i = 0; i < 6; ++i) ax[i] = v8[i];
      ax[6] = S;
      ax[7] = v8[7];
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int a = lane + 32 * k;
    if (a < d) { rec[a] = q[k]; rec[d + a] = p[k]; }
  }
  if (lane == 0) rec[2 * d] = S;
}

__global__ void __launch_bounds__(CHUNK_THREADS, 3)
k_rk4_chunk(EngDev E, PotDev P, double h, int nsteps, int traj0, int ntb, double2 *__restrict__ cm,
            const double *__restrict__ hd, ChunkLayout L) {
  constexpr int WM = 2, WN = CHUNK_WN, TPT = CHUNK_THREADS;
  extern __shared__ __align__(16) double smem[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int d = E.d, ldu = L.ldu, ldh = L.ldh, nb = L.nb, nc = L.nc, DK = L.dk, dp = (d + 1) & ~1;
  double *__restrict__ U = smem + L.off_U;
  double *__restrict__ V = smem + L.off_V;
  double *__restrict__ H = smem + L.off_H;
  double *vec = smem + L.off_vec;
  double *hdv = vec;                                   // [2][4][dp] Hessian diagonals, double buffered over steps
  double *csa = vec + 8 * dp, *cisa = vec + 9 * dp, *csb = vec + 10 * dp, *cisb = vec + 11 * dp;
  const int m0 = warp * WM * 8;
  const int fr = lane >> 2, fc = lane & 3;
  int bcol[WN];
#pragma unroll
  for (int j = 0; j < WN; ++j) bcol[j] = ((8 * j < 2 * nb) ? (8 * j + fr) : 0) ^ swz(fc);

  for (int i = t; i < d * ldh; i += TPT) H[i] = 0.0;
  if (t < d) {
    csa[t] = 0.5 * E.sgt[t];
    cisa[t] = 0.5 * E.isgt[t];
    csb[t] = E.sgi[t];
    cisb[t] = E.isgi[t];
  }
  PT_DECL
  const int nitems = ntb * nc;
  const int nhd = 4 * dp;                              // doubles per (step, trajectory) row of hd
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int tl = item / nc, chunk = item - tl * nc;
    const int traj = traj0 + tl;
    const int b0 = chunk * nb;
    const int nbc = (d - b0 < nb) ? d - b0 : nb;            // columns b of this chunk
    double *rec = E.rec + (size_t)traj * E.rs;
    __syncthreads();
    for (int idx = t; idx < DK * ldu; idx += TPT) U[idx] = 0.0;
    for (int idx = t; idx < d * ldu; idx += TPT) V[idx] = 0.0;
    for (int i = t; i < nhd; i += TPT) hdv[i] = hd[((size_t)0 * ntb + tl) * nhd + i];
    __syncthreads();
    // local column lc < nb: (Mqq, Mpq)[:, b0+lc];  lc >= nb: (Mqp, Mpp)[:, b0+lc-nb]; 8 loads in flight per thread
    {
      const int ntot = d * 2 * nbc;
      for (int base = 0; base < ntot; base += 8 * TPT) {
        double uv[8], vv[8];
        int la[8], lcs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int idx = base + k * TPT + t;
          uv[k] = vv[k] = 0.0;
          la[k] = -1;
          if (idx < ntot) {
            const int a = idx / (2 * nbc), r = idx - a * 2 * nbc;
            const int half = r / nbc, lb = r - half * nbc;
            const int gcol = half * d + b0 + lb;
            la[k] = a;
            lcs[k] = half * nb + lb;
            uv[k] = rec[E.qps + a * 2 * d + gcol];
            vv[k] = rec[E.qps + 2 * d * d + a * 2 * d + gcol];
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (la[k] >= 0) {
            U[la[k] * ldu + (lcs[k] ^ swz(la[k]))] = uv[k];
            V[la[k] * ldu + lcs[k]] = vv[k];
          }
      }
    }
    if (t < d) H[t * ldh + t] = hdv[t];
    __syncthreads();
    PT(0);

    for (int step = 0; step < nsteps; ++step) {
      const double *hcur = hdv + (step & 1) * nhd;
      // prefetch the Hessian diagonals of the next step (consumed after the last barrier of this step)
      double pre[2] = {0.0, 0.0};
      const bool has_next = step + 1 < nsteps;
      if (has_next) {
        const double *src = hd + ((size_t)(step + 1) * ntb + tl) * nhd;
        if (t < nhd) pre[0] = src[t];
        if (t + TPT < nhd) pre[1] = src[t + TPT];
      }
      double R1[WM][WN][2], R2[WM][WN][2];
#pragma unroll 1
      for (int s = 1; s <= 4; ++s) {
        // ---- phase A: acc = H_s U_s on the tensor pipe (software-pipelined fragment loads)
        double acc[WM][WN][2];
#pragma unroll
        for (int i = 0; i < WM; ++i)
#pragma unroll
          for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        {
          const int r0 = m0 + fr, r1 = m0 + 8 + fr;
          const double *Ha0 = H + (r0 < d ? r0 : 0) * ldh + fc, *Ha1 = H + (r1 < d ? r1 : 0) * ldh + fc;
          const bool v0 = r0 < d, v1 = r1 < d;
          const double *Bp = U + fc * ldu;
          const int ldu4 = 4 * ldu;
          double a0 = v0 ? Ha0[0] : 0.0, a1 = v1 ? Ha1[0] : 0.0, b[WN];
#pragma unroll
          for (int j = 0; j < WN; ++j) b[j] = Bp[bcol[j]];
          const int nk = DK >> 2;
#pragma unroll 5
          for (int k = 1; k <= nk; ++k) {
            double an0 = 0.0, an1 = 0.0, bn[WN];
            if (k < nk) {
              Ha0 += 4;
              Ha1 += 4;
              Bp += ldu4;
              an0 = v0 ? Ha0[0] : 0.0;
              an1 = v1 ? Ha1[0] : 0.0;
#pragma unroll
              for (int j = 0; j < WN; ++j) bn[j] = Bp[bcol[j]];
            }
#pragma unroll
            for (int j = 0; j < WN; ++j) {
              dmma884(acc[0][j][0], acc[0][j][1], a0, b[j]);
              dmma884(acc[1][j][0], acc[1][j][1], a1, b[j]);
            }
            a0 = an0;
            a1 = an1;
#pragma unroll
            for (int j = 0; j < WN; ++j) b[j] = bn[j];
          }
        }
        __syncthreads();
        PT(2);
        // ---- phase B: RK4 bookkeeping, next stage operand in place (kv = -acc); loads of a whole tile row are
        // issued before the arithmetic so that the shared-memory latency is paid once per row, not once per pair
        const double hh4 = 0.25 * h * h, hh2 = 0.5 * h * h, hh6 = h * h / 6.0, h2 = 0.5 * h, h6 = h / 6.0;
#pragma unroll
        for (int i = 0; i < WM; ++i) {
          const int a = m0 + 8 * i + fr;
          if (a < d) {
            const double ima = P.imass[a];
            const int sw = swz(a);
            double2 *urow = reinterpret_cast<double2 *>(U + a * ldu);
            double2 *vrow = reinterpret_cast<double2 *>(V + a * ldu);
            const bool need_v = (s != 2);
#pragma unroll
            for (int jg = 0; jg < WN; jg += 3) {
              constexpr int G = 3;
              double2 u[G], v[G];
#pragma unroll
              for (int jj = 0; jj < G; ++jj) {
                const int j = jg + jj, lc = 8 * j + 2 * fc;
                u[jj] = v[jj] = make_double2(0.0, 0.0);
                if (j < WN && lc < 2 * nb) {
                  u[jj] = urow[(lc ^ sw) >> 1];
                  if (need_v) v[jj] = vrow[lc >> 1];
                }
              }
#pragma unroll
              for (int jj = 0; jj < G; ++jj) {
                const int j = (jg + jj < WN) ? jg + jj : WN - 1, lc = 8 * (jg + jj) + 2 * fc;
                if (jg + jj < WN && lc < 2 * nb) {
                  const double k0 = -acc[i][j][0], k1 = -acc[i][j][1];
                  if (s == 1) {
                    R1[i][j][0] = k0; R1[i][j][1] = k1;
                    R2[i][j][0] = 0.0; R2[i][j][1] = 0.0;
                    u[jj].x += h2 * v[jj].x * ima;
                    u[jj].y += h2 * v[jj].y * ima;
                  } else if (s == 2) {
                    R2[i][j][0] = k0; R2[i][j][1] = k1;
                    u[jj].x += hh4 * R1[i][j][0] * ima;
                    u[jj].y += hh4 * R1[i][j][1] * ima;
                  } else if (s == 3) {
                    const double a1x = R1[i][j][0], a1y = R1[i][j][1], a2x = R2[i][j][0], a2y = R2[i][j][1];
                    u[jj].x += (h2 * v[jj].x + hh2 * a2x - hh4 * a1x) * ima;
                    u[jj].y += (h2 * v[jj].y + hh2 * a2y - hh4 * a1y) * ima;
                    R1[i][j][0] = hh6 * (a1x + a2x + k0) - hh2 * a2x;
                    R1[i][j][1] = hh6 * (a1y + a2y + k1) - hh2 * a2y;
                    R2[i][j][0] = a1x + 2.0 * a2x + 2.0 * k0;
                    R2[i][j][1] = a1y + 2.0 * a2y + 2.0 * k1;
                  } else {
                    u[jj].x += R1[i][j][0] * ima;
                    u[jj].y += R1[i][j][1] * ima;
                    v[jj].x += h6 * (R2[i][j][0] + k0);
                    v[jj].y += h6 * (R2[i][j][1] + k1);
                    vrow[lc >> 1] = v[jj];
                  }
                  urow[(lc ^ sw) >> 1] = u[jj];
                }
              }
            }
          }
        }
        if (t < d && s < 4) H[t * ldh + t] = hcur[s * dp + t];
        if (s == 4 && has_next) {
          double *hn = hdv + ((step + 1) & 1) * nhd;
          if (t < nhd) hn[t] = pre[0];
          if (t + TPT < nhd) hn[t + TPT] = pre[1];
          if (t < d) H[t * ldh + t] = pre[0];          // stage 1 of the next step (row 0 of the next buffer)
        }
        __syncthreads();
        PT(3);
      }
      // ---- this chunk's columns of the prefactor matrix (propagators.py:969-986 with diagonal width matrices)
      {
        double2 *out = cm + ((size_t)step * ntb + tl) * d * d;
        for (int idx = t; idx < d * nbc; idx += TPT) {
          const int a = idx / nbc, lb = idx - a * nbc, b = b0 + lb;
          const int sw = swz(a);
          const double mqq = U[a * ldu + (lb ^ sw)], mqp = U[a * ldu + ((nb + lb) ^ sw)];
          const double mpq = V[a * ldu + lb], mpp = V[a * ldu + nb + lb];
          const double sa = csa[a], isa = cisa[a], sb = csb[b], isb = cisb[b];
          out[a * d + b] = make_double2(sa * mqq * isb + isa * mpp * sb, -sa * mqp * sb + isa * mpq * isb);
        }
      }
      PT(4);
    }
    // ---- write back
    __syncthreads();
    for (int idx = t; idx < d * 2 * nbc; idx += TPT) {
      const int a = idx / (2 * nbc), r = idx - a * 2 * nbc;
      const int half = r / nbc, lb = r - half * nbc;
      const int lc = half * nb + lb, gcol = half * d + b0 + lb;
      rec[E.qps + a * 2 * d + gcol] = U[a * ldu + (lc ^ swz(a))];
      rec[E.qps + 2 * d * d + a * 2 * d + gcol] = V[a * ldu + lc];
    }
    PT(8);
  }
}

// one thread per trajectory of the batch, time steps in order; partials: (gridDim.x, nsteps_total, 5) rows of this
// batch's blocks (row stride nsteps_total, first step of this launch = step0)
__global__ void __launch_bounds__(128)
k_hk_finish(EngDev E, int traj0, int ntb, int nsteps, int step0, int nsteps_total, const double2 *__restrict__ det,
            const double *__restrict__ aux, double *__restrict__ partials) {
  __shared__ double red[4][5];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tl = blockIdx.x * blockDim.x + t;
  const bool live = tl < ntb;
  const int traj = traj0 + (live ? tl : 0);
  double2 c2 = E.c2[traj], cc = E.c[traj];
  double sign = E.sign[traj];
  const double2 wvi = E.wvi[traj];
  for (int step = 0; step < nsteps; ++step) {
    double v5[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (live) {
      const double2 dt = det[(size_t)step * ntb + tl];
      const double *ax = aux + ((size_t)step * ntb + tl) * 8;
      sign = track_sign(sign, c2, dt);
      c2 = dt;
      cc = csqrt_principal(dt);
      const double v6[6] = {ax[0], ax[1], ax[2], ax[3], ax[4], ax[5]};
      double2 ca, ki;
      corr_finish(E, v6, ax[6], cc, sign, wvi, ca, ki);
      v5[0] = ca.x; v5[1] = ca.y; v5[2] = ki.x; v5[3] = ki.y; v5[4] = ax[7];
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v5[i] += __shfl_xor_sync(0xffffffffu, v5[i], o);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 5; ++i) red[warp][i] = v5[i];
    }
    __syncthreads();
    if (t < 5) partials[((size_t)blockIdx.x * nsteps_total + step0 + step) * 5 + t] = (red[0][t] + red[1][t]) + (red[2][t] + red[3][t]);
  }
  if (live) {
    E.c2[traj] = c2;
    E.c[traj] = cc;
    E.sign[traj] = sign;
  }
}

static bool chunk_supported(const EngDev &E, const PotDev &P) {
  if (!E.diag || E.dr != E.d) return false;
  if (P.type != POT_MORSE && P.type != POT_NONHARMONIC) return false;
  if (E.d <= 32 || E.d > 62) return false;
  const ChunkLayout L = make_chunk_layout(E.d);
  return 2 * L.nb <= 8 * CHUNK_WN && E.d <= 64;
}

}  // namespace sc
