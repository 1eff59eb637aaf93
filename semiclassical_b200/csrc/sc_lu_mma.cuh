// sc_lu_mma.cuh -- batched complex LU determinant with the trailing updates on the FP64 tensor pipe
// (HK prefactor determinant, propagators.py:998-999; 32 < dr <= 64).
//
// Left-looking blocked LU with partial pivoting, dataflow-synchronised like lu_det_left (sc_lu.cuh), but a block
// column lives in registers in the DMMA *accumulator* layout and a published panel is applied to it as ONE complex
// rank-4 update on mma.sync.m8n8k4.f64 instead of 16 shuffle-driven rank-1 steps:
//
//     block column J (64 rows x 4 complex columns) = 8 accumulator tiles of 8 rows x 8 real columns
//                                                     (real column 2j = Re, 2j+1 = Im of complex column j)
//     panel K publishes  W_K = L_K inv(L_KK)   (64 x 4 complex, zero rows for retired rows, shared memory)
//     X   = rows p_0..p_3 (the pivot rows of panel K) of the up-to-date block column      (4 x 4 complex)
//     A_J <- A_J - W_K X :   [Re|Im] -= [Re W | Im W] [[Re X, Im X], [-Im X, Re X]]       (2 k-steps of 4 per tile)
//
// Retired rows never influence the determinant again, so their (U) values are simply left to rot and row tiles
// whose 8 rows are all retired are skipped.  The pivot rows are data dependent, and a dynamically indexed row of an
// accumulator tile is a dynamically indexed REGISTER; instead the warp keeps a mirror of its block column in the
// block's own (not yet published) panel slot in shared memory -- one STS.128 per live tile and update -- reads X
// from there with a computed address, and the same mirror is the transpose into the rows <-> lanes layout that the
// in-warp panel factorisation (lu_panel) needs.  After factoring, W is formed by a 4 x 4 back substitution and
// stored over the mirror in the A-fragment order (complex row-major [64][4], 16-byte XOR swizzle: conflict-free
// for the tile accesses and for the row accesses).
// Per (block, panel) pair: ~65 warp instructions (8 x {LDS.128, 2 DMMA, STS.128} + bookkeeping) instead of ~300.
#pragma once
#include "sc_lu.cuh"

namespace sc {


struct LuMmaFixed {
  double2 pv[LU_MAX_PANELS][4];
  int p[LU_MAX_PANELS][4];
  unsigned long long bar[LU_MAX_PANELS];
  unsigned long long dmask[LU_MAX_PANELS];   // rows retired once panel K is done (padding rows included)
  unsigned tmask[LU_MAX_PANELS];             // bit T: all 8 rows of row tile T are retired after panel K
  int inv[LU_MAX_PANELS];                    // inversions contributed by the pivots of panel K (permutation parity)
  int ready;
  int pad[3];
};
constexpr int LUM_PANEL_ELEMS = 64 * 4;   // double2 per panel slot
static inline size_t lum_smem_bytes(int dr) { return sizeof(LuMmaFixed) + (size_t)((dr + 3) / 4) * LUM_PANEL_ELEMS * sizeof(double2); }

__device__ __forceinline__ void lum_dmma(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ void lum_bar_init(LuMmaFixed *sh, int t) {
  if (t < LU_MAX_PANELS)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(&sh->bar[t]))));
  if (t == 0) sh->ready = 0;
}

__device__ __forceinline__ double2 lum_bcast(const double2 (&llo)[4], const double2 (&lhi)[4], int c, int p) {
  double2 v = (p >= 32) ? lhi[c] : llo[c];
  v.x = __shfl_sync(0xffffffffu, v.x, p & 31);
  v.y = __shfl_sync(0xffffffffu, v.y, p & 31);
  return v;
}
// a -= b c
__device__ __forceinline__ void lum_cfms(double2 &a, double2 b, double2 c) {
  a.x = fma(-b.x, c.x, fma(b.y, c.y, a.x));
  a.y = fma(-b.x, c.y, fma(-b.y, c.x, a.y));
}

// in-warp factorisation of panel J in the rows <-> lanes layout (lane l: rows l and l + 32); publishes pivots and,
// unless this is the last panel, W = L inv(L_KK) into WJ.  done: rows retired before this panel.
__device__ __forceinline__ void lum_panel(double2 (&lo)[4], double2 (&hi)[4], int ncol, unsigned long long done,
                                          unsigned long long real_rows, bool last, LuMmaFixed *sh, int J, double2 *WJ,
                                          int lane) {
  double2 llo[4], lhi[4];
  int pp[4];
  int inv = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    llo[c] = lhi[c] = make_double2(0.0, 0.0);
    pp[c] = 0;
    if (c < ncol) {
      const double mlo = lo[c].x * lo[c].x + lo[c].y * lo[c].y, mhi = hi[c].x * hi[c].x + hi[c].y * hi[c].y;
      unsigned key = 0u;
      if (!((done >> lane) & 1ull)) key = lu_key2(mlo, lane);
      if (!((done >> (lane + 32)) & 1ull)) key = max(key, lu_key2(mhi, lane + 32));
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
      inv += __popcll(((done & real_rows) >> p) >> 1);           // earlier pivots in rows below this one
      done |= 1ull << p;
      double2 flo = make_double2(lo[c].x * ix - lo[c].y * iy, lo[c].x * iy + lo[c].y * ix);
      double2 fhi = make_double2(hi[c].x * ix - hi[c].y * iy, hi[c].x * iy + hi[c].y * ix);
      if ((done >> lane) & 1ull) flo = make_double2(0.0, 0.0);
      if ((done >> (lane + 32)) & 1ull) fhi = make_double2(0.0, 0.0);
#pragma unroll
      for (int c2 = c + 1; c2 < 4; ++c2) lu_rank1(lo[c2], hi[c2], p, flo, fhi, true);
      llo[c] = flo;
      lhi[c] = fhi;
      pp[c] = p;
      if (lane == 0) {
        sh->p[J][c] = p;
        sh->pv[J][c] = make_double2(px, py);
      }
    } else if (lane == 0) {
      sh->p[J][c] = 0;
      sh->pv[J][c] = make_double2(1.0, 0.0);
    }
  }
  if (lane == 0) sh->inv[J] = inv;
  if (last) return;
  if (lane == 0) {
    unsigned long long f = done & (done >> 1);
    f &= f >> 2;
    f &= f >> 4;                                                 // bit 8T: all 8 rows of tile T retired
    unsigned tm = 0u;
#pragma unroll
    for (int T = 0; T < 8; ++T) tm |= static_cast<unsigned>((f >> (8 * T)) & 1ull) << T;
    sh->dmask[J] = done;
    sh->tmask[J] = tm;
  }
  // W L_KK = L with L_KK[i][c] = multiplier of pivot row p_i in column c (c < i), unit diagonal
  const double2 L10 = lum_bcast(llo, lhi, 0, pp[1]);
  const double2 L20 = lum_bcast(llo, lhi, 0, pp[2]), L21 = lum_bcast(llo, lhi, 1, pp[2]);
  const double2 L30 = lum_bcast(llo, lhi, 0, pp[3]), L31 = lum_bcast(llo, lhi, 1, pp[3]), L32 = lum_bcast(llo, lhi, 2, pp[3]);
  const int sw = (lane >> 1) & 3;
  auto finish = [&](double2(&l)[4], bool retired, int row) {
    lum_cfms(l[2], l[3], L32);
    lum_cfms(l[1], l[2], L21);
    lum_cfms(l[1], l[3], L31);
    lum_cfms(l[0], l[1], L10);
    lum_cfms(l[0], l[2], L20);
    lum_cfms(l[0], l[3], L30);
#pragma unroll
    for (int c = 0; c < 4; ++c) WJ[row * 4 + (c ^ sw)] = retired ? make_double2(0.0, 0.0) : l[c];
  };
  __syncwarp();   // every lane has read its rows of the transposed block from this slot
  finish(llo, (done >> lane) & 1ull, lane);
  finish(lhi, (done >> (lane + 32)) & 1ull, lane + 32);
}

// Same panel when one half of the rows (lanes' "lo" rows 0..31 or "hi" rows 32..63) is entirely retired -- the usual
// state of the second half of the factorisation when the pivots stay near the diagonal: one row per lane, half the
// instructions on the critical path.  HI: the live rows are lane + 32.
template <bool HI>
__device__ __forceinline__ void lum_panel1(double2 (&a)[4], int ncol, unsigned long long done, unsigned long long real_rows,
                                           bool last, LuMmaFixed *sh, int J, double2 *WJ, int lane) {
  const int row = lane + (HI ? 32 : 0);
  bool dead = (done >> row) & 1ull;
  const unsigned real_half = static_cast<unsigned>(HI ? real_rows >> 32 : real_rows);
  const int above = HI ? 0 : __popc(static_cast<unsigned>(real_rows >> 32));   // retired real rows of the other half below p
  int inv = 0;
  double2 l[4];
  int pp[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    l[c] = make_double2(0.0, 0.0);
    pp[c] = row;
    if (c < ncol) {
      const double m = a[c].x * a[c].x + a[c].y * a[c].y;
      const unsigned key = dead ? 0u : lu_key2(m, row);
      const double rr = fast_rcp(m == 0.0 ? 1.0 : m);
      const unsigned kk = __reduce_max_sync(0xffffffffu, key);
      const int p = static_cast<int>(kk & 63u), src = p & 31;
      inv += __popc(((__ballot_sync(0xffffffffu, dead) & real_half) >> src) >> 1) + above;
      double px = a[c].x, py = a[c].y;
      double ix = px * rr, iy = -py * rr;
      px = __shfl_sync(0xffffffffu, px, src);
      py = __shfl_sync(0xffffffffu, py, src);
      ix = __shfl_sync(0xffffffffu, ix, src);
      iy = __shfl_sync(0xffffffffu, iy, src);
      dead = dead || (row == p);
      double2 f = make_double2(a[c].x * ix - a[c].y * iy, a[c].x * iy + a[c].y * ix);
      if (dead) f = make_double2(0.0, 0.0);
#pragma unroll
      for (int c2 = c + 1; c2 < 4; ++c2) {
        const double ux = __shfl_sync(0xffffffffu, a[c2].x, src), uy = __shfl_sync(0xffffffffu, a[c2].y, src);
        lum_cfms(a[c2], f, make_double2(ux, uy));
      }
      l[c] = f;
      pp[c] = p;
      if (lane == 0) {
        sh->p[J][c] = p;
        sh->pv[J][c] = make_double2(px, py);
      }
    } else if (lane == 0) {
      sh->p[J][c] = 0;
      sh->pv[J][c] = make_double2(1.0, 0.0);
    }
  }
  if (lane == 0) sh->inv[J] = inv;
  if (last) return;
  const unsigned live = __ballot_sync(0xffffffffu, !dead);
  if (lane == 0) {
    const unsigned long long dn = HI ? ((static_cast<unsigned long long>(~live) << 32) | 0xffffffffull)
                                     : (static_cast<unsigned long long>(~live) | 0xffffffff00000000ull);
    unsigned long long f = dn & (dn >> 1);
    f &= f >> 2;
    f &= f >> 4;
    unsigned tm = 0u;
#pragma unroll
    for (int T = 0; T < 8; ++T) tm |= static_cast<unsigned>((f >> (8 * T)) & 1ull) << T;
    sh->dmask[J] = dn;
    sh->tmask[J] = tm;
  }
  auto bc = [&](int c, int p) {
    return make_double2(__shfl_sync(0xffffffffu, l[c].x, p & 31), __shfl_sync(0xffffffffu, l[c].y, p & 31));
  };
  const double2 L10 = bc(0, pp[1]);
  const double2 L20 = bc(0, pp[2]), L21 = bc(1, pp[2]);
  const double2 L30 = bc(0, pp[3]), L31 = bc(1, pp[3]), L32 = bc(2, pp[3]);
  lum_cfms(l[2], l[3], L32);
  lum_cfms(l[1], l[2], L21);
  lum_cfms(l[1], l[3], L31);
  lum_cfms(l[0], l[1], L10);
  lum_cfms(l[0], l[2], L20);
  lum_cfms(l[0], l[3], L30);
  const int sw = (lane >> 1) & 3, other = lane + (HI ? 0 : 32);
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    WJ[row * 4 + (c ^ sw)] = dead ? make_double2(0.0, 0.0) : l[c];
    WJ[other * 4 + c] = make_double2(0.0, 0.0);
  }
}

// block column J of A in the accumulator layout (zero padding); A == nullptr: nothing to load
__device__ __forceinline__ void lum_load_block(const double2 *__restrict__ A, int ld, int dr, int J, int g, int j, double2 (&v)[8]) {
  const int col = 4 * J + j;
#pragma unroll
  for (int T = 0; T < 8; ++T) {
    const int r = 8 * T + g;
    v[T] = make_double2(0.0, 0.0);
    if (A != nullptr && col < dr && r < dr) v[T] = A[(size_t)col * ld + r];
  }
}

// A: element (LU row r, LU column c) at A[c * ld + r]; 32 < dr <= 64.  W: nblocks panel slots.  Determinant on warp 0.
template <int NW>
__device__ __forceinline__ double2 lum_det(const double2 *__restrict__ A, const double2 *__restrict__ Anext, double2 (&nxt)[8],
                                           int ld, int dr, LuMmaFixed *sh, double2 *W, int base, unsigned parity, int w,
                                           int lane) {
  const int nblocks = (dr + 3) >> 2;
  const int g = lane >> 2, j = lane & 3;
  const int swl = (lane & ~3) | (j ^ ((g >> 1) & 3));          // swizzled position of (row g of a tile, column j)
  const int part = g & 1, jj = lane >> 3;                       // B fragment: k = j, n = g = 2 jj + part
  const unsigned long long pad_rows = dr < 64 ? (~0ull << dr) : 0ull;
  int known = 0;
  for (int J = w; J < nblocks; J += NW) {
    double2 *WJ = W + (size_t)J * LUM_PANEL_ELEMS;               // mirror of this block column until it is factored
    double acc[8][2];
#pragma unroll
    for (int T = 0; T < 8; ++T) {                                // loaded while the previous block was being factored
      acc[T][0] = nxt[T].x;
      acc[T][1] = nxt[T].y;
      WJ[T * 32 + swl] = nxt[T];
    }
    unsigned long long done = pad_rows;
    // next block of this warp (or its first block of the CTA's next matrix): the loads are issued when the warp first
    // has to wait for a panel (it is ahead of the factorisation front, i.e. about to be on the critical path, and the
    // address arithmetic costs nothing there), else right before its own panel; they fly during the factorisation
    bool prefetched = false;
    const bool same = J + NW < nblocks;
#pragma unroll 1
    for (int K = 0; K < J; ++K) {
      if (known <= K) {
        known = flow_peek(&sh->ready) - base;
        if (known <= K) {
          if (!prefetched) {
            lum_load_block(same ? A : Anext, ld, dr, same ? J + NW : w, g, j, nxt);
            prefetched = true;
          }
          flow_bar_wait(&sh->bar[K], parity);
          known = K + 1;
        }
      }
      const int pk = sh->p[K][j];
      const unsigned full = sh->tmask[K];
      const double2 *Wk = W + (size_t)K * LUM_PANEL_ELEMS + swl;
      double2 wv[8];
#pragma unroll
      for (int T = 0; T < 8; ++T)
        if (!((full >> T) & 1u)) wv[T] = Wk[T * 32];             // warp-uniform: some row of this tile is still active
      __syncwarp();                                              // mirror stores of the previous update are visible
      const double2 x = WJ[pk * 4 + (jj ^ ((pk >> 1) & 3))];     // X[k = j][jj]
      __syncwarp();                                              // ... and read before this update overwrites them
      const double b0 = part ? -x.y : -x.x, b1 = part ? -x.x : x.y;
      // the two k-steps of a tile are dependent (26 cycles); issue all first k-steps, then all second ones
#pragma unroll
      for (int T = 0; T < 8; ++T)
        if (!((full >> T) & 1u)) lum_dmma(acc[T][0], acc[T][1], wv[T].x, b0);
#pragma unroll
      for (int T = 0; T < 8; ++T)
        if (!((full >> T) & 1u)) {
          lum_dmma(acc[T][0], acc[T][1], wv[T].y, b1);
          WJ[T * 32 + swl] = make_double2(acc[T][0], acc[T][1]);
        }
    }
    if (J > 0) done = sh->dmask[J - 1];
    if (!prefetched) lum_load_block(same ? A : Anext, ld, dr, same ? J + NW : w, g, j, nxt);
    // the mirror is the transpose into the rows <-> lanes layout: factor, publish
    __syncwarp();
    {
      const int sw = (lane >> 1) & 3, ncol = min(4, dr - 4 * J);
      const bool last = J + 1 == nblocks;
      double2 lo[4], hi[4];
      if (static_cast<unsigned>(done) == 0xffffffffu) {          // warp-uniform
#pragma unroll
        for (int c = 0; c < 4; ++c) hi[c] = WJ[(lane + 32) * 4 + (c ^ sw)];
        lum_panel1<true>(hi, ncol, done, ~pad_rows, last, sh, J, WJ, lane);
      } else if (static_cast<unsigned>(done >> 32) == 0xffffffffu) {
#pragma unroll
        for (int c = 0; c < 4; ++c) lo[c] = WJ[lane * 4 + (c ^ sw)];
        lum_panel1<false>(lo, ncol, done, ~pad_rows, last, sh, J, WJ, lane);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          lo[c] = WJ[lane * 4 + (c ^ sw)];
          hi[c] = WJ[(lane + 32) * 4 + (c ^ sw)];
        }
        lum_panel(lo, hi, min(4, dr - 4 * J), done, ~pad_rows, last, sh, J, WJ, lane);
      }
    }
    __syncwarp();
    if (lane == 0) {
      flow_publish(&sh->ready, base + J + 1);
      flow_bar_arrive(&sh->bar[J]);
    }
  }
  double2 det = make_double2(1.0, 0.0);
  if (w == 0) {
    flow_bar_wait(&sh->bar[nblocks - 1], parity);
    int inv = lane < nblocks ? sh->inv[lane] : 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k = lane + 32 * half;
      if (k < dr) det = cmul(det, sh->pv[k >> 2][k & 3]);
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

template <int NW, int OCC>
__global__ void __launch_bounds__(32 * NW, OCC)
k_lu_mma(const double2 *__restrict__ mats, int dr, int nmat, double2 *__restrict__ det_out) {
  extern __shared__ __align__(16) unsigned char lu_smem[];
  LuMmaFixed *sh = reinterpret_cast<LuMmaFixed *>(lu_smem);
  double2 *W = reinterpret_cast<double2 *>(lu_smem + sizeof(LuMmaFixed));
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int nblocks = (dr + 3) >> 2;
  lum_bar_init(sh, t);
  int base = 0;
  unsigned parity = 0;
  double2 nxt[8];
  lum_load_block(blockIdx.x < nmat ? mats + (size_t)blockIdx.x * dr * dr : nullptr, dr, dr, w, lane >> 2, lane & 3, nxt);
  for (int mat = blockIdx.x; mat < nmat; mat += gridDim.x) {
    __syncthreads();   // panel slots and barriers of the previous matrix are no longer in use
    const int matn = mat + gridDim.x;
    const double2 det = lum_det<NW>(mats + (size_t)mat * dr * dr, matn < nmat ? mats + (size_t)matn * dr * dr : nullptr, nxt, dr,
                                    dr, sh, W, base, parity, w, lane);
    base += nblocks;
    parity ^= 1u;
    if (t == 0) det_out[mat] = det;
  }
}

}  // namespace sc
