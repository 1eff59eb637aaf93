// sc_lu.cuh -- batched complex LU determinant, register resident, blocked (HK prefactor, propagators.py:998-999).
//
// The prefactor matrices of many (trajectory, time step) pairs are independent, so the determinant is computed by
// a throughput kernel: one CTA of NW warps per matrix at a time, several CTAs per SM.  Layout of one matrix inside
// the CTA (never touches shared memory):
//     lane l owns the rows l ("lo") and l + 32 ("hi"); columns are grouped in blocks of 4, block J belongs to warp
//     J % NW (slot J / NW): each thread holds 2 x 4 x NBLK complex elements.
// Right-looking blocked elimination with implicit partial pivoting (rows are never moved; retired rows get zero
// multipliers, the permutation parity comes from popcounts of the retired-row mask):
//     * the warp that owns block K factors its 4-column panel entirely in-warp (redux.sync pivot search on a packed
//       (magnitude, row) key, shuffles for the pivot row) and publishes the 4 multiplier columns + pivots
//     * ONE CTA barrier per block; the owner of block K+1 updates that block first and factors it while the other
//       warps are still applying panel K to their trailing columns (look-ahead), double-buffered panels
//     * trailing update of a column: 4 sequential rank-1 steps, pivot-row element by shuffle from lane p % 32
// Since det A = det A^T the caller may load either orientation (whichever is coalesced).
#pragma once
#include "sc_device.cuh"

namespace sc {

struct LuPanel {
  double2 f[4][64];     // multipliers of the 4 panel columns (0 for retired rows)
  double2 pv[4];        // pivot values
  int p[4];             // pivot rows
  int pad[4];
};

__device__ __forceinline__ unsigned lu_key2(double m, int row) {
  // top bits of the squared magnitude, 6 low bits = row; +64 so that an exact zero still beats "no candidate"
  return (static_cast<unsigned>(__double2hiint(m)) & ~63u) + 64u + static_cast<unsigned>(row);
}

// 1/x for x > 0: hardware approximation (2^-23) + two Newton steps (relative error ~1e-14, no slow path); the
// multipliers do not need a correctly rounded reciprocal
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * fma(-x, r, 2.0);
  r = r * fma(-x, r, 2.0);
  return r;
}

// rank-1 step on one column: u = a[p] (shuffle), a -= f u
__device__ __forceinline__ void lu_rank1(double2 &lo, double2 &hi, int p, double2 flo, double2 fhi, bool use_hi) {
  const bool ph = p >= 32;
  double ux = ph ? hi.x : lo.x, uy = ph ? hi.y : lo.y;
  ux = __shfl_sync(0xffffffffu, ux, p & 31);
  uy = __shfl_sync(0xffffffffu, uy, p & 31);
  lo.x = fma(-flo.x, ux, fma(flo.y, uy, lo.x));
  lo.y = fma(-flo.x, uy, fma(-flo.y, ux, lo.y));
  if (use_hi) {
    hi.x = fma(-fhi.x, ux, fma(fhi.y, uy, hi.x));
    hi.y = fma(-fhi.x, uy, fma(-fhi.y, ux, hi.y));
  }
}

// in-warp factorisation of one 4-column panel (columns c0..c0+3 of the matrix; only the first ncol are real)
__device__ __forceinline__ void lu_panel(double2 (&lo)[4], double2 (&hi)[4], int ncol, unsigned long long &done,
                                         bool use_hi, LuPanel *out, int lane) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c < ncol) {
      const double mlo = lo[c].x * lo[c].x + lo[c].y * lo[c].y, mhi = hi[c].x * hi[c].x + hi[c].y * hi[c].y;
      unsigned key = 0u;
      if (!((done >> lane) & 1ull)) key = lu_key2(mlo, lane);
      if (use_hi && !((done >> (lane + 32)) & 1ull)) key = max(key, lu_key2(mhi, lane + 32));
      // reciprocals of both candidates while the redux is in flight (exact zeros must not hit the slow path)
      const double rlo = fast_rcp(mlo == 0.0 ? 1.0 : mlo), rhi = fast_rcp(mhi == 0.0 ? 1.0 : mhi);
      const unsigned kk = __reduce_max_sync(0xffffffffu, key);
      const int p = static_cast<int>(kk & 63u);
      const bool ph = p >= 32;
      double px = ph ? hi[c].x : lo[c].x, py = ph ? hi[c].y : lo[c].y;
      const double rr = ph ? rhi : rlo;
      double ix = px * rr, iy = -py * rr;
      px = __shfl_sync(0xffffffffu, px, p & 31);
      py = __shfl_sync(0xffffffffu, py, p & 31);
      ix = __shfl_sync(0xffffffffu, ix, p & 31);
      iy = __shfl_sync(0xffffffffu, iy, p & 31);
      done |= 1ull << p;
      double2 flo = make_double2(lo[c].x * ix - lo[c].y * iy, lo[c].x * iy + lo[c].y * ix);
      double2 fhi = make_double2(hi[c].x * ix - hi[c].y * iy, hi[c].x * iy + hi[c].y * ix);
      if ((done >> lane) & 1ull) flo = make_double2(0.0, 0.0);
      if (!use_hi || ((done >> (lane + 32)) & 1ull)) fhi = make_double2(0.0, 0.0);
#pragma unroll
      for (int c2 = c + 1; c2 < 4; ++c2) lu_rank1(lo[c2], hi[c2], p, flo, fhi, use_hi);
      out->f[c][lane] = flo;
      out->f[c][lane + 32] = fhi;
      if (lane == 0) {
        out->p[c] = p;
        out->pv[c] = make_double2(px, py);
      }
    }
  }
}

// lo[s][c], hi[s][c]: element (row lane / lane+32, column 4 (w + NW s) + c); entries outside dr x dr must be zero.
// sh: 2 panels of shared memory.  BAR_ID: named barrier used by the 32 NW threads of this matrix.
// Result valid on every participating thread.
template <int NW, int NBLK, int BAR_ID>
__device__ __forceinline__ double2 lu_det_blk(double2 (&lo)[NBLK][4], double2 (&hi)[NBLK][4], int dr, LuPanel *sh, int w,
                                              int lane) {
  const bool use_hi = dr > 32;
  const int nblocks = (dr + 3) >> 2;
  unsigned long long done = 0ull;       // retired rows (identical on every thread)
  if (w == 0) {
    unsigned long long dn = 0ull;
    lu_panel(lo[0], hi[0], min(4, dr), dn, use_hi, &sh[0], lane);
  }
  lu_bar<BAR_ID, 32 * NW>();
  double2 det = make_double2(1.0, 0.0);
  int inversions = 0;
#pragma unroll
  for (int s = 0; s < NBLK; ++s) {
#pragma unroll 1
    for (int ww = 0; ww < NW; ++ww) {
      const int K = s * NW + ww;              // panel being applied
      if (K >= nblocks) break;
      const LuPanel *P = &sh[K & 1];
      const int ncol = min(4, dr - 4 * K);
      int p[4];
      double2 flo[4], fhi[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        p[c] = P->p[c];
        flo[c] = P->f[c][lane];
        fhi[c] = P->f[c][lane + 32];
        if (c < ncol) {
          det = cmul(det, P->pv[c]);
          inversions += __popcll(done >> p[c]);   // earlier pivots with a larger row index
          done |= 1ull << p[c];
        } else {
          flo[c] = fhi[c] = make_double2(0.0, 0.0);
          p[c] = 0;
        }
      }
      if (K + 1 < nblocks) {
        auto apply = [&](double2(&l)[4], double2(&h)[4]) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) lu_rank1(l[j], h[j], p[c], flo[c], fhi[c], use_hi);
          }
        };
        // look-ahead: the owner of block K+1 brings it up to date, factors it and publishes it first
        const bool wrap = (ww + 1 == NW);
        const int wn = wrap ? 0 : ww + 1;
        if (w == wn) {
          unsigned long long dn = done;
          if (wrap) {
            if (s + 1 < NBLK) {
              apply(lo[s + 1 < NBLK ? s + 1 : s], hi[s + 1 < NBLK ? s + 1 : s]);
              lu_panel(lo[s + 1 < NBLK ? s + 1 : s], hi[s + 1 < NBLK ? s + 1 : s], min(4, dr - 4 * (K + 1)), dn, use_hi,
                       &sh[(K + 1) & 1], lane);
            }
          } else {
            apply(lo[s], hi[s]);
            lu_panel(lo[s], hi[s], min(4, dr - 4 * (K + 1)), dn, use_hi, &sh[(K + 1) & 1], lane);
          }
        }
        // remaining trailing blocks of this warp: slot s if its block lies beyond K+1, all later slots
        if (w > ww + 1) apply(lo[s], hi[s]);
#pragma unroll
        for (int s2 = s + 1; s2 < NBLK; ++s2) {
          if (!(wrap && w == 0 && s2 == s + 1)) apply(lo[s2], hi[s2]);
        }
        lu_bar<BAR_ID, 32 * NW>();
      }
    }
  }
  if (inversions & 1) { det.x = -det.x; det.y = -det.y; }
  return det;
}

// ------------------------------------------------------------------ dataflow variant ---------
// Same factorisation without CTA-wide barriers: every panel has its own slot in shared memory and a release/acquire
// counter says how many panels are published, so a warp waits only for the panel it needs.  The critical path is
// then  [apply panel K to block K+1] -> [factor panel K+1] -> publish, executed by the owner of block K+1, while
// all other warps run their trailing updates whenever their inputs are there.
constexpr int LU_MAX_PANELS = 16;
struct LuFlow {
  LuPanel panel[LU_MAX_PANELS];
  unsigned long long bar[LU_MAX_PANELS];   // one mbarrier per panel (arrival count 1), one phase per matrix
  int ready;
  int pad[3];
};

__device__ __forceinline__ void flow_bar_init(LuFlow *sh, int t) {
  if (t < LU_MAX_PANELS)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(&sh->bar[t]))));
  if (t == 0) sh->ready = 0;
}
__device__ __forceinline__ void flow_bar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))) : "memory");
}
__device__ __forceinline__ void flow_bar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LU_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LU_DONE;\n"
      "bra LU_WAIT;\n"
      "LU_DONE:\n"
      "}\n" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ int flow_peek(int *flag) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(static_cast<unsigned>(__cvta_generic_to_shared(flag))) : "memory");
  return v;
}

__device__ __forceinline__ void flow_publish(int *flag, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(flag))), "r"(v) : "memory");
}
__device__ __forceinline__ void flow_wait(int *flag, int want) {
  int v;
  do {
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(static_cast<unsigned>(__cvta_generic_to_shared(flag))) : "memory");
  } while (v < want);
}

// ------------------------------------------------------------------ left-looking dataflow variant ----------
// Each warp brings ONE 4-column block at a time up to date: it applies the published panels 0..J-1 in order (waiting
// on the release/acquire counter only when it is ahead of the factorisation front), factors its block in-warp and
// publishes panel J.  Only 8 complex elements per thread are live, the multipliers of all panels stay in shared
// memory (LuFlow, 66 KB) -> three matrices per SM are in flight and every warp's instruction stream is
// load-multipliers / 16 rank-1 steps per (block, panel) pair.  Critical path per block: one panel application +
// one in-warp panel factorisation.
// A: matrix in global memory, element (LU row r, LU column c) at A[c * ld + r] (the transpose of a row-major
// matrix; det is the same).  Returns the determinant on warp 0 (all lanes); other warps return garbage.
template <int NW>
__device__ __forceinline__ double2 lu_det_left(const double2 *__restrict__ A, int ld, int dr, LuFlow *sh, int base,
                                               unsigned parity, int w, int lane) {
  const bool use_hi = dr > 32;
  const int nblocks = (dr + 3) >> 2;
  int known = 0;                       // panels of this matrix known to be published
  for (int J = w; J < nblocks; J += NW) {
    double2 lo[4], hi[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = 4 * J + c;
      lo[c] = hi[c] = make_double2(0.0, 0.0);
      if (col < dr) {
        if (lane < dr) lo[c] = A[(size_t)col * ld + lane];
        if (lane + 32 < dr) hi[c] = A[(size_t)col * ld + lane + 32];
      }
    }
    unsigned long long done = 0ull;
#pragma unroll 1
    for (int K = 0; K < J; ++K) {
      if (known <= K) {
        known = flow_peek(&sh->ready) - base;
        if (known <= K) {               // ahead of the factorisation front: sleep on the panel's mbarrier
          flow_bar_wait(&sh->bar[K], parity);
          known = K + 1;
        }
      }
      const LuPanel *P = &sh->panel[K];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int p = P->p[c];
        const double2 flo = P->f[c][lane], fhi = P->f[c][lane + 32];
        done |= 1ull << p;
#pragma unroll
        for (int j = 0; j < 4; ++j) lu_rank1(lo[j], hi[j], p, flo, fhi, use_hi);
      }
    }
    lu_panel(lo, hi, min(4, dr - 4 * J), done, use_hi, &sh->panel[J], lane);
    __syncwarp();
    if (lane == 0) {
      flow_publish(&sh->ready, base + J + 1);
      flow_bar_arrive(&sh->bar[J]);
    }
  }
  double2 det = make_double2(1.0, 0.0);
  if (w == 0) {
    flow_bar_wait(&sh->bar[nblocks - 1], parity);
    int inv = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k = lane + 32 * half;
      if (k < dr) {
        const int pk = sh->panel[k >> 2].p[k & 3];
        det = cmul(det, sh->panel[k >> 2].pv[k & 3]);
        for (int k2 = 0; k2 < k; ++k2) inv += (sh->panel[k2 >> 2].p[k2 & 3] > pk) ? 1 : 0;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ox = __shfl_xor_sync(0xffffffffu, det.x, o), oy = __shfl_xor_sync(0xffffffffu, det.y, o);
      det = cmul(det, make_double2(ox, oy));
      inv += __shfl_xor_sync(0xffffffffu, inv, o);
    }
    if (inv & 1) { det.x = -det.x; det.y = -det.y; }
  }
  return det;
}

}  // namespace sc
