// sc_chunk.cuh -- column pipeline of the Herman-Kluk step for large d with diagonal width matrices (the headline path).
//
// The 2d columns of the monodromy blocks are independent linear ODEs driven by the same Hessians
// (propagators.py:342-357: d/dt [Mqq;Mpq][:,b] and d/dt [Mqp;Mpp][:,b] only involve column b), and with diagonal
// Gamma column b of the prefactor matrix (propagators.py:969-986) needs exactly column b of the four blocks.  The
// step is therefore split into throughput kernels:
//
//   k_qp_path     (q, p, S) of the separable model for K steps, one warp per trajectory: Hessian diagonals of every
//                 stage, overlap / NAC partial sums, action, T+V.
//   k_rk4_wcols   work item = (trajectory, tile of 4 columns b) owned by ONE warp for K steps: the warp's 60 x 8 slab
//                 of U = [Mqq|Mqp] is the B operand of H U_s (mma.sync.m8n8k4.f64) and the set of elements it updates;
//                 RK4 accumulators in registers, stage operand formed IN PLACE:
//                     U_2 = U + h/2 V/m,  U_3 = U_2 + h^2/4 kv_1/m,  U_4 = U_3 + [h/2 V + h^2/2 kv_2 - h^2/4 kv_1]/m,
//                     U'  = U_4 + [h^2/6 (kv_1+kv_2+kv_3) - h^2/2 kv_2]/m,   V' = V + h/6 (kv_1+2kv_2+2kv_3+kv_4)
//                 (kv_s = -H_s U_s; algebraically the classical RK4 of propagators.py:114-119).  H_s = H0 + diag(h_s):
//                 the dense base H0 (shared memory, zero for the separable models served here) is multiplied in
//                 full, the stage diagonal is added to the A fragment of the diagonal tile.  Every step the warp
//                 writes its 4 columns of the complex prefactor matrix to a global scratch.
//   k_lu_mma      (sc_lu_mma.cuh, sc_lu_batch.cuh) determinants of all (step, trajectory) matrices of the batch.
//   k_hk_finish   per trajectory, in time order: sqrt branch tracking (propagators.py:1045-1047), contributions to
//                 C_auto and k_ic (propagators.py:784-911), deterministic per-block partial sums.
// Layout of a slab in shared memory: row-major [row][8]; B-fragment loads (4 rows x 8 columns = 256 contiguous
// bytes) and the 128-bit owner accesses (8 rows x 64 bytes) are conflict free without swizzling.
#pragma once
#include "sc_mma.cuh"

namespace sc {

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

template <int S>
__device__ __forceinline__ void cols_phase_b(double2 &u, double2 &v, double ima, double h, double k0, double k1, double (&R1)[2],
                                             double (&R2)[2]) {
  const double hh4 = 0.25 * h * h, hh2 = 0.5 * h * h, hh6 = h * h / 6.0, h2 = 0.5 * h, h6 = h / 6.0;
  if (S == 1) {
    R1[0] = k0; R1[1] = k1;
    u.x = fma(h2 * ima, v.x, u.x);
    u.y = fma(h2 * ima, v.y, u.y);
  } else if (S == 2) {
    R2[0] = k0; R2[1] = k1;
    u.x = fma(hh4 * ima, R1[0], u.x);
    u.y = fma(hh4 * ima, R1[1], u.y);
  } else if (S == 3) {
    const double a1x = R1[0], a1y = R1[1], a2x = R2[0], a2y = R2[1];
    u.x += (h2 * v.x + hh2 * a2x - hh4 * a1x) * ima;
    u.y += (h2 * v.y + hh2 * a2y - hh4 * a1y) * ima;
    R1[0] = hh6 * (a1x + a2x + k0) - hh2 * a2x;
    R1[1] = hh6 * (a1y + a2y + k1) - hh2 * a2y;
    R2[0] = a1x + 2.0 * a2x + 2.0 * k0;
    R2[1] = a1y + 2.0 * a2y + 2.0 * k1;
  } else {
    u.x = fma(R1[0], ima, u.x);
    u.y = fma(R1[1], ima, u.y);
    v.x = fma(h6, R2[0] + k0, v.x);
    v.y = fma(h6, R2[1] + k1, v.y);
  }
}

// leading dimension of the Hessian base for NK k-steps: >= 4 NK with ld mod 16 in {4, 12} (conflict-free A fragments)
__host__ __device__ constexpr int cols_ldh(int nk) { return (4 * nk) % 16 == 4 || (4 * nk) % 16 == 12 ? 4 * nk : ((4 * nk + 4) % 16 == 4 || (4 * nk + 4) % 16 == 12 ? 4 * nk + 4 : 4 * nk + 8); }

// ------------------------------------------------------------------ warp-independent variant ----------------
// Same arithmetic as k_rk4_cols with the work item shrunk to (trajectory, ONE tile of 4 columns b) and owned by a
// single warp for all time steps of the launch: no CTA barrier after the set-up (k_rk4_cols: one per time step,
// 10 % of the warp time) and no idle warp when ceil(d / 4) is not a multiple of 4 (d = 60: 15 tiles, one warp slot
// in 16 was empty).  The Hessian diagonals of a step (4 stages x d) live in a warp-private 4 x dp buffer that is
// refilled by cp.async while the step is still running: stages 1-3 of step n + 1 when stage 4 of step n starts
// (their region is dead by then), stage 4 right after the stage-4 MMA phase.
struct WColsLayout {
  int nt, ntw, nitem, ldh, dk;                      // column tiles per trajectory, tiles per warp, warp items per trajectory
  int off_H, off_c, off_W, slab, wstride, off_hdw, total; // doubles; per-warp region: ntw x (U slab, V slab), hd buffer
};
__host__ __device__ inline WColsLayout make_wcols_layout(int d, int ntw) {
  WColsLayout L;
  L.nt = (d + 3) / 4;
  L.ntw = ntw;
  L.nitem = (L.nt + ntw - 1) / ntw;
  L.dk = (d + 3) & ~3;
  L.ldh = cols_ldh(L.dk / 4);
  const int dp = (d + 1) & ~1;
  int o = 0;
  L.off_H = o; o += d * L.ldh;
  L.off_c = o; o += 6 * dp;                         // sa, isa, sb, isb, 1/m (+ zero padding up to 8 MT rows)
  o = (o + 1) & ~1;
  L.off_W = o;
  L.slab = (L.dk + d) * 8;
  L.off_hdw = ntw * L.slab;
  L.wstride = L.off_hdw + 4 * dp;
  o += 4 * L.wstride;
  L.total = (o + 1) & ~1;
  return L;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// NTW = column tiles per warp.  NTW = 2: every A fragment (Hessian) feeds two DMMAs -- the fragment loads were the
// first stall of the one-tile kernel (shared-memory pipe 60 % busy) -- at the price of 2 x 48 accumulator / RK4
// registers per thread, i.e. two CTAs per SM instead of three.
template <int NK, int NTW>
__global__ void __launch_bounds__(128, (NTW == 1 ? 3 : 2))
k_rk4_wcols(EngDev E, PotDev P, double h, int nsteps, int traj0, int ntb, double2 *__restrict__ cm,
            const double *__restrict__ hd, WColsLayout L, unsigned long long *__restrict__ queue) {
  constexpr int MT = (NK + 1) / 2;                    // 8-row tiles
  constexpr int LDH = cols_ldh(NK), DK = 4 * NK;
  extern __shared__ __align__(16) double smem[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nthr = blockDim.x;
  const int d = E.d, dp = (d + 1) & ~1;
  double *__restrict__ Wreg = smem + L.off_W + warp * L.wstride;
  double *__restrict__ hdw = Wreg + L.off_hdw;        // [4][dp], this warp's current time step
  const double *csa = smem + L.off_c, *cisa = csa + dp, *csb = cisa + dp, *cisb = csb + dp;
  const int fr = lane >> 2, fc = lane & 3;
  const bool pe = (fr == fc), po = (fr == fc + 4);    // lanes holding a diagonal element in even / odd k-steps
  const bool last_ok = (8 * (MT - 1) + fr) < d;       // only the last row tile can stick out of the matrix
  const double *__restrict__ Hfr = smem + L.off_H + (last_ok ? fr : 0) * LDH + fc;
  const double *__restrict__ Hfr0 = smem + L.off_H + fr * LDH + fc;
  const double *__restrict__ Ub = Wreg + fc * 8 + fr;                                       // + w * slab
  double2 *__restrict__ Uo = reinterpret_cast<double2 *>(Wreg + fr * 8 + 2 * fc);           // + w * slab / 2
  double2 *__restrict__ Vo = reinterpret_cast<double2 *>(Wreg + DK * 8 + fr * 8 + 2 * fc);
  const int slab = L.slab, slab2 = L.slab / 2;
  const double *cim = cisb + dp + fr;                   // 1 / m of row 8 i + fr at cim[8 i] (zero beyond d)
  double ima[MT];                                       // NTW = 1 keeps them in registers (measured: 3.6 % faster)
#pragma unroll
  for (int i = 0; i < MT; ++i) ima[i] = (8 * i + fr < d) ? P.imass[8 * i + fr] : 0.0;

  // dense base of the Hessian: off-diagonal part (identically zero for the separable models served here)
  for (int i = t; i < d * LDH; i += nthr) smem[L.off_H + i] = 0.0;
  if (t < 2 * dp) {
    double *c = smem + L.off_c;
    if (t < d) {
      c[t] = 0.5 * E.sgt[t];
      c[dp + t] = 0.5 * E.isgt[t];
      c[2 * dp + t] = E.sgi[t];
      c[3 * dp + t] = E.isgi[t];
    }
    c[4 * dp + t] = t < d ? P.imass[t] : 0.0;
  }
  __syncthreads();                                      // the only CTA barrier
  const int nt = L.nt, nitem = L.nitem;
  const long long nitems = (long long)ntb * nitem;
  const int nhd = 4 * dp;
  const int c16_123 = 3 * dp / 2, c16_4 = dp / 2;       // 16-byte pieces of the stage 1-3 rows / of the stage-4 row
  // items are handed out dynamically (one atomic per ~100 us item): no tail of warps with one item more than others
  auto next_item = [&]() {
    unsigned long long v = 0ull;
    if (lane == 0) v = atomicAdd(queue, 1ull);
    return static_cast<long long>(__shfl_sync(0xffffffffu, v, 0));
  };
  for (long long item = next_item(); item < nitems; item = next_item()) {
    const int tl = static_cast<int>(item / nitem), it = static_cast<int>(item - (long long)tl * nitem);
    const int traj = traj0 + tl;
    int b[NTW];
    bool bok[NTW];
#pragma unroll
    for (int w = 0; w < NTW; ++w) {
      b[w] = 4 * (NTW * it + w) + fc;                       // the column b this thread's elements of tile w belong to
      bok[w] = b[w] < d;
    }
    const bool two = NTW > 1 && (NTW * it + 1 < nt);        // warp-uniform: the second tile exists
    double *rec = E.rec + (size_t)traj * E.rs;
    __syncwarp();                                           // the previous item of this warp is finished
    {
      const double *src = hd + (size_t)tl * nhd;
      for (int i = lane; i < 2 * dp; i += 32) cp_async16(hdw + 2 * i, src + 2 * i);
    }
    // ---- load the slabs: row a = 8 i + fr, element pair (2 fc, 2 fc + 1) = (q-half, p-half) of column b
#pragma unroll
    for (int w = 0; w < NTW; ++w) {
      double2 u[MT], v[MT];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int a = 8 * i + fr;
        u[i] = v[i] = make_double2(0.0, 0.0);
        if (a < d && bok[w]) {
          const double *ru = rec + E.qps + (size_t)a * 2 * d, *rv = ru + 2 * d * d;
          u[i] = make_double2(ru[b[w]], ru[d + b[w]]);
          v[i] = make_double2(rv[b[w]], rv[d + b[w]]);
        }
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int a = 8 * i + fr;
        if (a < DK) Uo[w * slab2 + i * 32] = u[i];
        if (a < d) Vo[w * slab2 + i * 32] = v[i];
      }
    }
    cp_async_wait_all();
    __syncwarp();

    for (int step = 0; step < nsteps; ++step) {
      const bool has_next = step + 1 < nsteps;
      const double *hnext = hd + ((size_t)(step + 1) * ntb + tl) * nhd;
      double R1[NTW][MT][2], R2[NTW][MT][2];
      double2 *outb = cm + ((size_t)step * ntb + tl) * d * d + (size_t)fr * d;
#pragma unroll 1
      for (int s = 1; s <= 4; ++s) {
        const double *__restrict__ hsf = hdw + (s - 1) * dp + fr;
        if (s == 4 && has_next) {                            // rows of stages 1-3 are dead: refill them for step + 1
          for (int i = lane; i < c16_123; i += 32) cp_async16(hdw + 2 * i, hnext + 2 * i);
        }
        // ---- H_s U_s on the tensor pipe: 8-row tiles x this warp's 8 (16) columns
        double acc[NTW][MT][2];
#pragma unroll
        for (int w = 0; w < NTW; ++w)
#pragma unroll
          for (int i = 0; i < MT; ++i) acc[w][i][0] = acc[w][i][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) {
          double bf[NTW];
#pragma unroll
          for (int w = 0; w < NTW; ++w) bf[w] = Ub[w * slab + kk * 32];
          double af[MT];
#pragma unroll
          for (int i = 0; i < MT - 1; ++i) af[i] = Hfr0[i * 8 * LDH + 4 * kk];
          af[MT - 1] = last_ok ? Hfr[(MT - 1) * 8 * LDH + 4 * kk] : 0.0;
          {
            const int id = kk >> 1;                          // tile that contains the diagonal of this k-step
            const bool on = ((kk & 1) ? po : pe) && (id < MT - 1 || last_ok);
            const double hv = hsf[(8 * id < 8 * (MT - 1) || last_ok) ? 8 * id : 0];
            if (on) af[id] += hv;
          }
#pragma unroll
          for (int i = 0; i < MT; ++i) dmma884(acc[0][i][0], acc[0][i][1], af[i], bf[0]);
          if (NTW > 1 && two) {
#pragma unroll
            for (int i = 0; i < MT; ++i) dmma884(acc[NTW - 1][i][0], acc[NTW - 1][i][1], af[i], bf[NTW - 1]);
          }
        }
        __syncwarp();
        if (s == 4 && has_next) {                            // the stage-4 row has been consumed by the MMA phase
          for (int i = lane; i < c16_4; i += 32) cp_async16(hdw + 3 * dp + 2 * i, hnext + 3 * dp + 2 * i);
        }
        // ---- RK4 bookkeeping on the warp's own slabs, stage operand in place
#pragma unroll
        for (int w = 0; w < NTW; ++w) {
          if (w > 0 && !two) break;
          double2 *Uw = Uo + w * slab2, *Vw = Vo + w * slab2;
          if (s == 1) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
              if (i < MT - 1 || last_ok) {
                double2 u = Uw[i * 32], v = Vw[i * 32];
                cols_phase_b<1>(u, v, (NTW == 1 ? ima[i] : cim[8 * i]), h, -acc[w][i][0], -acc[w][i][1], R1[w][i], R2[w][i]);
                Uw[i * 32] = u;
              }
          } else if (s == 2) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
              if (i < MT - 1 || last_ok) {
                double2 u = Uw[i * 32], v = make_double2(0.0, 0.0);
                cols_phase_b<2>(u, v, (NTW == 1 ? ima[i] : cim[8 * i]), h, -acc[w][i][0], -acc[w][i][1], R1[w][i], R2[w][i]);
                Uw[i * 32] = u;
              }
          } else if (s == 3) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
              if (i < MT - 1 || last_ok) {
                double2 u = Uw[i * 32], v = Vw[i * 32];
                cols_phase_b<3>(u, v, (NTW == 1 ? ima[i] : cim[8 * i]), h, -acc[w][i][0], -acc[w][i][1], R1[w][i], R2[w][i]);
                Uw[i * 32] = u;
              }
          } else {
            const double sb = bok[w] ? csb[b[w]] : 0.0, isb = bok[w] ? cisb[b[w]] : 0.0;
            double2 *out = outb + (bok[w] ? b[w] : 0);
#pragma unroll
            for (int i = 0; i < MT; ++i)
              if (i < MT - 1 || last_ok) {
                double2 u = Uw[i * 32], v = Vw[i * 32];
                cols_phase_b<4>(u, v, (NTW == 1 ? ima[i] : cim[8 * i]), h, -acc[w][i][0], -acc[w][i][1], R1[w][i], R2[w][i]);
                Uw[i * 32] = u;
                Vw[i * 32] = v;
                // prefactor-matrix element (propagators.py:969-986, diagonal width matrices) straight from registers:
                // u = (Mqq, Mqp)[a][b], v = (Mpq, Mpp)[a][b]
                if (bok[w]) {
                  const double sa = csa[8 * i + fr], isa = cisa[8 * i + fr];
                  out[(size_t)i * 8 * d] = make_double2(sa * u.x * isb + isa * v.y * sb, -sa * u.y * sb + isa * v.x * isb);
                }
              }
          }
        }
        __syncwarp();
      }
      cp_async_wait_all();
      __syncwarp();
    }
    // ---- write back
#pragma unroll
    for (int w = 0; w < NTW; ++w) {
      if (bok[w]) {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const int a = 8 * i + fr;
          if (a < d) {
            const double2 u = Uo[w * slab2 + i * 32], v = Vo[w * slab2 + i * 32];
            double *ru = rec + E.qps + (size_t)a * 2 * d, *rv = ru + 2 * d * d;
            ru[b[w]] = u.x; ru[d + b[w]] = u.y;
            rv[b[w]] = v.x; rv[d + b[w]] = v.y;
          }
        }
      }
    }
  }
}

template <int NK, int NTW>
static cudaError_t launch_wcols_t(int grid, size_t smem, const EngDev &E, const PotDev &P, double h, int nsteps, int traj0, int ntb,
                                  double2 *cm, const double *hd, const WColsLayout &L, unsigned long long *queue, cudaStream_t st) {
  cudaError_t ce = cudaFuncSetAttribute(k_rk4_wcols<NK, NTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce != cudaSuccess) return ce;
  // full shared-memory carve-out: CTAs of the LU kernel can become resident next to this kernel's without an SM re-configuration
  ce = cudaFuncSetAttribute(k_rk4_wcols<NK, NTW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (ce != cudaSuccess) return ce;
  ce = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st);   // work queue of this launch
  if (ce != cudaSuccess) return ce;
  k_rk4_wcols<NK, NTW><<<grid, 128, smem, st>>>(E, P, h, nsteps, traj0, ntb, cm, hd, L, queue);
  return cudaGetLastError();
}

static cudaError_t launch_wcols(int grid, const EngDev &E, const PotDev &P, double h, int nsteps, int traj0, int ntb, double2 *cm,
                                const double *hd, const WColsLayout &L, unsigned long long *queue, cudaStream_t st) {
  const size_t smem = sizeof(double) * (size_t)L.total;
  switch (L.dk / 4) {
#define SC_WCOLS_CASE(N)                                                                                              \
  case N:                                                                                                             \
    return launch_wcols_t<N, 1>(grid, smem, E, P, h, nsteps, traj0, ntb, cm, hd, L, queue, st);
    SC_WCOLS_CASE(5) SC_WCOLS_CASE(6) SC_WCOLS_CASE(7) SC_WCOLS_CASE(8) SC_WCOLS_CASE(9) SC_WCOLS_CASE(10) SC_WCOLS_CASE(11) SC_WCOLS_CASE(12) SC_WCOLS_CASE(13) SC_WCOLS_CASE(14) SC_WCOLS_CASE(15)
    SC_WCOLS_CASE(16)
#undef SC_WCOLS_CASE
    default: return cudaErrorInvalidValue;
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
  // 17 <= d <= 64 (the work queue of k_rk4_wcols hands out (trajectory, column tile) items, so small systems keep all warps
  // busy; measured against the dense pipeline on the AS model: +26 % at d = 17, +20 % at 24, +24 % at 32)
  int dmin = 17;
  if (const char *s = getenv("SC_CHUNK_DMIN")) dmin = atoi(s) > 16 ? atoi(s) : 17;
  return E.d >= dmin && E.d <= 64;
}

}  // namespace sc
