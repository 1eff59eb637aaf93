// sc_stream.cuh -- dense column pipeline: the Herman-Kluk step for ANY potential (per-trajectory dense Hessians) and ANY
// width matrices (dense, rank deficient), 13 <= d <= 96 (d <= 64 for per-trajectory Hessians).
//
// Same structural facts as sc_chunk.cuh: the 2d columns of the monodromy blocks are independent linear ODEs driven by the
// same Hessians (propagators.py:342-357), and the (q, p, S) path does not depend on the monodromy matrices.  The step is
// split into throughput kernels over a window of trajectories x K time steps:
//
//   k_path_*      (q, p, S) for K steps, one warp per trajectory; per step: q, p, S, T+V of the 4th stage point.  Potentials
//                 with a trajectory-dependent Hessian also write the four stage Hessians H_s of every step (d x LDH fp64, the
//                 shared-memory image the matrix kernel wants) to a global stream; the harmonic molecule's Hessian is
//                 constant (potentials.py:581-593) and streams from ONE padded copy.
//   k_aux_terms   overlap / NAC partial sums of every (step, trajectory) from q, p (propagators.py:230-237, 868-909).
//   k_rk4_stream  one CTA = one trajectory at a time, warp w owns NTW tiles of 4 columns b of [Mqq|Mqp] and [Mpq|Mpp]
//                 (60 x 8 slabs, warp private, resident for K steps).  The A operands of all warps -- H_1..H_4 of every
//                 step, and for dense width matrices the left factors L1, L2 of the prefactor (propagators.py:969-994) as
//                 "stages 5 and 6" -- arrive through a ring of NS shared-memory slots filled by cp.async.bulk (TMA engine,
//                 one mbarrier per slot).  There is no CTA barrier: a warp waits for slot g, multiplies
//                 (mma.sync.m8n8k4.f64), releases the slot (the LAST warp to release issues the copy of stage g + NS into
//                 it), then does its RK4 bookkeeping in place on its own slabs (cols_phase_b, sc_chunk.cuh).
//                 Diagonal widths: the 4 columns of the prefactor matrix straight from registers.  Dense widths: the
//                 C fragments of L1 [Mqq|Mqp] and L2 [Mpq|Mpp] go to a scratch in the A-fragment order of k_rmult.
//   k_rmult       Re C = 1/2 (L1 Mqq) R1 + 1/2 (L2 Mpp) R2,  Im C = 1/2 (L2 Mpq) R1 - 1/2 (L1 Mqp) R2: all (step, trajectory)
//                 matrices stacked into one tall GEMM with the constant B operand [R1; R2] in shared memory.
//   k_lu_mma / k_lu_batch, k_hk_finish  as in the column pipeline of sc_chunk.cuh.
//
// Algorithmic flops per trajectory-step: 16 d^3 (RK4) + 8 d' d^2 (left factors) + 8 d'^2 d (right factors) + 8/3 d'^3 (LU).
#pragma once
#include <cstdint>

#include "sc_chunk.cuh"
#include "sc_gdml.cuh"

namespace sc {

struct StreamLayout {
  int nt, ntw, nwarp, ngroups, dk, ldh, hsz, ns, mtr;   // tiles, tiles per warp, warps per CTA, CTAs (tile groups) per trajectory, ...
  int off_ring, off_c, off_W, slab, wstride, total;   // doubles
};

// nwarp = 0: one CTA holds all tiles of a trajectory; else nwarp warps per CTA and ceil(nt / (nwarp ntw)) CTAs per trajectory
__host__ __device__ inline StreamLayout make_stream_layout(int d, int dr, int ntw, int ns, int nwarp = 0) {
  StreamLayout L;
  L.nt = (d + 3) / 4;
  L.ntw = ntw;
  L.nwarp = nwarp > 0 ? nwarp : (L.nt + ntw - 1) / ntw;
  L.ngroups = (L.nt + L.nwarp * ntw - 1) / (L.nwarp * ntw);
  L.dk = (d + 3) & ~3;
  L.ldh = cols_ldh(L.dk / 4);
  L.hsz = d * L.ldh;
  L.ns = ns;
  L.mtr = (dr + 7) / 8;
  const int dp = (d + 1) & ~1;
  int o = 0;
  L.off_ring = o; o += ns * L.hsz;
  L.off_c = o; o += 6 * dp + 8;                   // sa, isa, sb, isb, 1/m (+ zero padding up to 8 MT rows)
  o = (o + 1) & ~1;
  L.off_W = o;
  L.slab = L.dk * 8;                               // one slab (U or V) of one tile: dk rows x 8 columns
  L.wstride = 2 * ntw * L.slab;
  o += L.nwarp * L.wstride;
  L.total = (o + 1) & ~1;
  return L;
}

struct StreamArgs {
  const double *hs;              // Hessian stream: matrix of (step, stage s, trajectory tl) at hs + ((step 4 + s) ntb + tl) hsz; or constant
  int hs_const;                  // 1: every stage reads the same matrix at hs
  const double *L1p, *L2p;       // dense widths: left factors padded to d x ldh (zero rows >= dr), else null
  double2 *cm;                   // diagonal widths: prefactor matrices (step, tl, d, d)
  double *T;                     // dense widths: (step, tl, row tile, column tile, plane, 32 lanes x 2) fragments of L1 U', L2 V'
  int skip_rk4;                  // 1: prefactor matrices of the records as they are (one "step", no propagation, no write back)
};

// ------------------------------------------------------------------ matrix kernel
constexpr int stream_max_threads(int nk, int ntw) { return nk <= 16 ? 32 * ((nk + ntw - 1) / ntw) : 256; }
// small systems leave most of the shared memory free: several CTAs (trajectories) per SM, registers capped accordingly
constexpr int stream_min_ctas(int nk) { return nk <= 6 ? 5 : nk <= 8 ? 4 : nk <= 10 ? 2 : 1; }

template <int NK, int NTW>
__global__ void __launch_bounds__(stream_max_threads(NK, NTW), stream_min_ctas(NK))
k_rk4_stream(EngDev E, PotDev P, double h, int nsteps, int traj0, int ntb, StreamArgs A, StreamLayout L) {
  constexpr int MT = (NK + 1) / 2;                    // 8-row tiles
  constexpr int LDH = cols_ldh(NK), DK = 4 * NK;
  constexpr int SLAB = DK * 8, SLAB2 = SLAB / 2;
  constexpr int MAXS = 4;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t full[MAXS];
  __shared__ int cnt[MAXS];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int d = E.d, dp = (d + 1) & ~1, dr = E.dr;
  // NK <= 16: one CTA holds all tiles of a trajectory (compile-time warp count, no ragged groups); above: tile groups
  constexpr bool GROUPS = NK > 16;
  constexpr int NWARP = (NK + NTW - 1) / NTW;
  const int NS = L.ns, hsz = L.hsz;
  const int nwarps = GROUPS ? (int)(blockDim.x >> 5) : NWARP, ngroups = GROUPS ? L.ngroups : 1;
  const bool dense = A.T != nullptr;
  const int nrk = A.skip_rk4 ? 0 : 4;
  const int nstg = nrk + (dense ? 2 : 0);
  double *ring = smem + L.off_ring;
  double *__restrict__ Wreg = smem + L.off_W + warp * L.wstride;     // tile w: U slab at w 2 SLAB, V slab at w 2 SLAB + SLAB
  const double *csa = smem + L.off_c, *cisa = csa + dp, *csb = cisa + dp, *cisb = csb + dp;
  const double *cim = cisb + dp;                        // 1 / m (zero beyond d, up to 8 MT)
  const int fr = lane >> 2, fc = lane & 3;
  const bool last_ok = (8 * (MT - 1) + fr) < d;         // only the last row tile can stick out of the matrix
  const int frl = last_ok ? fr : 0;
  const double *__restrict__ Ub = Wreg + fc * 8 + fr;                                       // + w 2 SLAB (+ SLAB for V)
  double2 *__restrict__ Uo = reinterpret_cast<double2 *>(Wreg + fr * 8 + 2 * fc);           // + w SLAB (double2 units: 2 SLAB2)
  const uint32_t hbytes = (uint32_t)(hsz * sizeof(double));

  // ---- set-up: constants, barriers, the first NS stream matrices
  if (t < MAXS) cnt[t] = 0;
  if (t == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = t; i < 8 * MT + 4 * dp; i += blockDim.x) {
    double *c = smem + L.off_c;
    if (i < 4 * dp) {
      const int k = i / dp, a = i - k * dp;
      double v = 0.0;
      if (a < d && E.diag) v = (k == 0) ? 0.5 * E.sgt[a] : (k == 1) ? 0.5 * E.isgt[a] : (k == 2) ? E.sgi[a] : E.isgi[a];
      c[i] = v;
    } else {
      const int a = i - 4 * dp;
      c[4 * dp + a] = a < d ? P.imass[a] : 0.0;
    }
  }
  const long long nitems = (long long)ntb * ngroups;       // item = (trajectory, tile group)
  const int n_iter = (int)((nitems - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x);
  const int per_traj = nsteps * nstg;
  const long long G = (long long)n_iter * per_traj;
  const int per_traj_d = per_traj > 0 ? per_traj : 1;
  auto issue = [&](long long g) {                       // one thread: copy of stage g into its slot
    if (g >= G) return;
    const int it = (int)(g / per_traj_d), rem = (int)(g - (long long)it * per_traj_d);
    const int step = rem / nstg, sidx = rem - step * nstg + (4 - nrk);
    const int tl = GROUPS ? (int)(((long long)blockIdx.x + (long long)it * gridDim.x) / ngroups) : (int)blockIdx.x + it * (int)gridDim.x;
    const double *src;
    if (sidx < 4) src = A.hs_const ? A.hs : A.hs + ((size_t)(step * 4 + sidx) * ntb + tl) * hsz;
    else src = (sidx == 4) ? A.L1p : A.L2p;
    const int slot = (int)(g % NS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&full[slot], hbytes);
    bulk_g2s(ring + (size_t)slot * hsz, src, hbytes, &full[slot]);
  };
  __syncthreads();                                      // the only CTA barrier
  if (t == 0)
    for (int i = 0; i < NS; ++i) issue(i);

  long long g = 0;
  // wait for stage g, acc = A_g B (B: this warp's U or V slabs), release the slot
  // mt_lim: row tiles beyond it are all padding (left prefactor factors with d' <= 8 (MT - 1)): skipped, warp-uniform
  auto mma_stage = [&](const double *__restrict__ Bb, bool valid, bool two, double (&acc)[NTW][MT][2], int mt_lim) {
    const int slot = (int)(g % NS);
    mbar_wait(&full[slot], (uint32_t)((g / NS) & 1));
    if (valid) {
    const double *__restrict__ Hs = ring + (size_t)slot * hsz + fc;
    const double *__restrict__ Hfr0 = Hs + fr * LDH, *__restrict__ Hfrl = Hs + frl * LDH;
#pragma unroll
    for (int w = 0; w < NTW; ++w)
#pragma unroll
      for (int i = 0; i < MT; ++i) acc[w][i][0] = acc[w][i][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < NK; ++kk) {
      double bf[NTW];
#pragma unroll
      for (int w = 0; w < NTW; ++w) bf[w] = Bb[w * 2 * SLAB + kk * 32];
      double af[MT];
#pragma unroll
      for (int i = 0; i < MT - 1; ++i) af[i] = Hfr0[i * 8 * LDH + 4 * kk];
      af[MT - 1] = (last_ok && MT - 1 < mt_lim) ? Hfrl[(MT - 1) * 8 * LDH + 4 * kk] : 0.0;
#pragma unroll
      for (int i = 0; i < MT - 1; ++i) dmma884(acc[0][i][0], acc[0][i][1], af[i], bf[0]);
      if (MT - 1 < mt_lim) dmma884(acc[0][MT - 1][0], acc[0][MT - 1][1], af[MT - 1], bf[0]);
      if (NTW > 1 && two) {
#pragma unroll
        for (int i = 0; i < MT - 1; ++i) dmma884(acc[NTW - 1][i][0], acc[NTW - 1][i][1], af[i], bf[NTW - 1]);
        if (MT - 1 < mt_lim) dmma884(acc[NTW - 1][MT - 1][0], acc[NTW - 1][MT - 1][1], af[MT - 1], bf[NTW - 1]);
      }
    }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const int old = atomicAdd(&cnt[slot], 1);
      if (old == nwarps - 1) {                           // every warp is done with the slot: refill it
        cnt[slot] = 0;
        __threadfence_block();
        issue(g + NS);
      }
    }
    ++g;
  };

  for (int it = 0; it < n_iter; ++it) {
    int tl, grp = 0;
    if (GROUPS) {
      const long long item = (long long)blockIdx.x + (long long)it * gridDim.x;
      tl = (int)(item / ngroups);
      grp = (int)(item - (long long)tl * ngroups);
    } else {
      tl = (int)blockIdx.x + it * (int)gridDim.x;
    }
    const int traj = traj0 + tl;
    const int tile0 = (grp * nwarps + warp) * NTW;          // first column tile of this warp
    const bool valid = GROUPS ? tile0 < NK : true;          // warp-uniform: the last group of a trajectory may be ragged
    int b[NTW];
    bool bok[NTW];
#pragma unroll
    for (int w = 0; w < NTW; ++w) {
      b[w] = 4 * (tile0 + w) + fc;                          // the column b this thread's elements of tile w belong to
      bok[w] = valid && b[w] < d;
    }
    const bool two = NTW > 1 && (tile0 + 1 < NK);           // warp-uniform: the second tile exists
    double *rec = E.rec + (size_t)traj * E.rs;
    __syncwarp();
    // ---- load the slabs: row a = 8 i + fr, element pair (2 fc, 2 fc + 1) = (q-half, p-half) of column b
#pragma unroll
    for (int w = 0; w < NTW; ++w) {
      double2 u[MT], v[MT];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int a = 8 * i + fr;
        u[i] = v[i] = make_double2(0.0, 0.0);
        if (a < d && bok[w] && (w == 0 || two)) {
          const double *ru = rec + E.qps + (size_t)a * 2 * d, *rv = ru + 2 * d * d;
          u[i] = make_double2(ru[b[w]], ru[d + b[w]]);
          v[i] = make_double2(rv[b[w]], rv[d + b[w]]);
        }
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int a = 8 * i + fr;
        if (a < DK) {
          Uo[w * 2 * SLAB2 + i * 32] = u[i];
          Uo[w * 2 * SLAB2 + SLAB2 + i * 32] = v[i];
        }
      }
    }
    __syncwarp();

    for (int step = 0; step < nsteps; ++step) {
      double R1[NTW][MT][2], R2[NTW][MT][2];
      double acc[NTW][MT][2];
      const size_t mat = (size_t)step * ntb + tl;
#pragma unroll 1
      for (int s = 1; s <= nrk; ++s) {
        mma_stage(Ub, valid, two, acc, MT);
        // ---- RK4 bookkeeping on the warp's own slabs, stage operand in place
#pragma unroll
        for (int w = 0; w < NTW; ++w) {
          if (!valid || (w > 0 && !two)) break;
          double2 *Uw = Uo + w * 2 * SLAB2, *Vw = Uw + SLAB2;
          if (s == 1) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
              if (i < MT - 1 || last_ok) {
                double2 u = Uw[i * 32], v = Vw[i * 32];
                cols_phase_b<1>(u, v, cim[8 * i + fr], h, -acc[w][i][0], -acc[w][i][1], R1[w][i], R2[w][i]);
                Uw[i * 32] = u;
              }
          } else if (s == 2) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
              if (i < MT - 1 || last_ok) {
                double2 u = Uw[i * 32], v = make_double2(0.0, 0.0);
                cols_phase_b<2>(u, v, cim[8 * i + fr], h, -acc[w][i][0], -acc[w][i][1], R1[w][i], R2[w][i]);
                Uw[i * 32] = u;
              }
          } else if (s == 3) {
#pragma unroll
            for (int i = 0; i < MT; ++i)
              if (i < MT - 1 || last_ok) {
                double2 u = Uw[i * 32], v = Vw[i * 32];
                cols_phase_b<3>(u, v, cim[8 * i + fr], h, -acc[w][i][0], -acc[w][i][1], R1[w][i], R2[w][i]);
                Uw[i * 32] = u;
              }
          } else {
#pragma unroll
            for (int i = 0; i < MT; ++i)
              if (i < MT - 1 || last_ok) {
                double2 u = Uw[i * 32], v = Vw[i * 32];
                cols_phase_b<4>(u, v, cim[8 * i + fr], h, -acc[w][i][0], -acc[w][i][1], R1[w][i], R2[w][i]);
                Uw[i * 32] = u;
                Vw[i * 32] = v;
              }
          }
        }
        __syncwarp();
      }
      if (!dense) {
        // ---- diagonal widths: the warp's 4 columns of the prefactor matrix (propagators.py:969-986) from its slabs:
        // u = (Mqq, Mqp)[a][b], v = (Mpq, Mpp)[a][b]
#pragma unroll
        for (int w = 0; w < NTW; ++w) {
          if ((w > 0 && !two) || !bok[w]) continue;
          const double2 *Uw = Uo + w * 2 * SLAB2, *Vw = Uw + SLAB2;
          const double sb = csb[b[w]], isb = cisb[b[w]];
          double2 *out = A.cm + mat * d * d + (size_t)fr * d + b[w];
#pragma unroll
          for (int i = 0; i < MT; ++i)
            if (i < MT - 1 || last_ok) {
              const double2 u = Uw[i * 32], v = Vw[i * 32];
              const double sa = csa[8 * i + fr], isa = cisa[8 * i + fr];
              out[(size_t)i * 8 * d] = make_double2(sa * u.x * isb + isa * v.y * sb, -sa * u.y * sb + isa * v.x * isb);
            }
        }
      } else {
        // ---- stages 5, 6: L1 [Mqq|Mqp] and L2 [Mpq|Mpp] of this warp's columns -> fragment scratch of k_rmult
        const int mtr = L.mtr;
#pragma unroll 1
        for (int pl = 0; pl < 2; ++pl) {
          mma_stage(Ub + pl * SLAB, valid, two, acc, mtr);
#pragma unroll
          for (int w = 0; w < NTW; ++w) {
            if (!valid || (w > 0 && !two)) break;
            const int tile = tile0 + w;
            double2 *Tp = reinterpret_cast<double2 *>(A.T) + ((mat * mtr) * NK + tile) * 64 + pl * 32 + lane;
#pragma unroll
            for (int i = 0; i < MT; ++i)
              if (i < mtr) Tp[(size_t)i * NK * 64] = make_double2(acc[w][i][0], acc[w][i][1]);
          }
        }
        __syncwarp();
      }
    }
    // ---- write back
    if (A.skip_rk4) continue;
#pragma unroll
    for (int w = 0; w < NTW; ++w) {
      if (bok[w] && (w == 0 || two)) {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const int a = 8 * i + fr;
          if (a < d) {
            const double2 u = Uo[w * 2 * SLAB2 + i * 32], v = Uo[w * 2 * SLAB2 + SLAB2 + i * 32];
            double *ru = rec + E.qps + (size_t)a * 2 * d, *rv = ru + 2 * d * d;
            ru[b[w]] = u.x; ru[d + b[w]] = u.y;
            rv[b[w]] = v.x; rv[d + b[w]] = v.y;
          }
        }
      }
    }
  }
  (void)dr;
}

// work = CTAs' worth of items (trajectories x tile groups); the grid is one wave: SMs x resident CTAs of this instantiation
template <int NK, int NTW>
static cudaError_t launch_stream_t(long long work, int sm, size_t smem, const EngDev &E, const PotDev &P, double h, int nsteps, int traj0,
                                   int ntb, const StreamArgs &A, const StreamLayout &L, cudaStream_t st) {
  static size_t occ_smem = ~(size_t)0;
  static int occ = 1;
  if (occ_smem != smem) {
    cudaError_t ce = cudaFuncSetAttribute(k_rk4_stream<NK, NTW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return ce;
    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rk4_stream<NK, NTW>, 32 * L.nwarp, smem);
    if (ce != cudaSuccess) return ce;
    if (occ < 1) occ = 1;
    occ_smem = smem;
  }
  const long long grid = work < (long long)sm * occ ? work : (long long)sm * occ;
  k_rk4_stream<NK, NTW><<<(int)(grid < 1 ? 1 : grid), 32 * L.nwarp, smem, st>>>(E, P, h, nsteps, traj0, ntb, A, L);
  return cudaGetLastError();
}

static cudaError_t launch_stream(long long work, int sm, const EngDev &E, const PotDev &P, double h, int nsteps, int traj0, int ntb,
                                 const StreamArgs &A, const StreamLayout &L, cudaStream_t st) {
  const size_t smem = sizeof(double) * (size_t)L.total;
  switch (L.dk / 4) {
#define SC_STREAM_CASE(N) \
  case N: return launch_stream_t<N, 2>(work, sm, smem, E, P, h, nsteps, traj0, ntb, A, L, st);
    SC_STREAM_CASE(4) SC_STREAM_CASE(5) SC_STREAM_CASE(6) SC_STREAM_CASE(7) SC_STREAM_CASE(8) SC_STREAM_CASE(9) SC_STREAM_CASE(10)
    SC_STREAM_CASE(11) SC_STREAM_CASE(12) SC_STREAM_CASE(13) SC_STREAM_CASE(14) SC_STREAM_CASE(15) SC_STREAM_CASE(16)
#undef SC_STREAM_CASE
    // 64 < d <= 96: one tile per warp, several CTAs (tile groups) per trajectory, each streaming the Hessians itself
#define SC_STREAM_CASE1(N) \
  case N: return launch_stream_t<N, 1>(work, sm, smem, E, P, h, nsteps, traj0, ntb, A, L, st);
    SC_STREAM_CASE1(17) SC_STREAM_CASE1(18) SC_STREAM_CASE1(19) SC_STREAM_CASE1(20) SC_STREAM_CASE1(21) SC_STREAM_CASE1(22)
    SC_STREAM_CASE1(23) SC_STREAM_CASE1(24)
#undef SC_STREAM_CASE1
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------ right factors of the dense prefactor
// item = (matrix, 8-row tile i): C[8 i .. 8 i + 7][:] from the fragment scratch of k_rk4_stream.  The K dimension runs over
// the column b of the monodromy blocks in the order the producer warps owned them (tile t, lane component fc), so the
// producer's C fragments ARE this kernel's A fragments.  R1, R2 live in shared memory as slabs [n tile][b][8] (the B
// fragment of one k-step is 256 contiguous bytes).  Output: cm[mat][r][n] complex (the LU takes the transpose, det A^T = det A).
struct RmultLayout { int nta, dk, total; };
__host__ __device__ inline RmultLayout make_rmult_layout(int d, int dr) {
  RmultLayout L;
  L.nta = (dr + 7) / 8;
  L.dk = (d + 3) & ~3;
  L.total = 2 * L.nta * L.dk * 8;
  return L;
}

template <int NTA>
__global__ void __launch_bounds__(256)
k_rmult(EngDev E, long long nitems, const double *__restrict__ T, double2 *__restrict__ cm, RmultLayout L) {
  extern __shared__ __align__(16) double rsm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
  const int d = E.d, dr = E.dr, dk = L.dk, nt = dk / 4, mtr = (dr + 7) / 8;
  const int fr = lane >> 2, fc = lane & 3;
  double *R1s = rsm, *R2s = rsm + NTA * dk * 8;
  for (int i = t; i < NTA * dk * 8; i += blockDim.x) {
    const int n = i / (dk * 8), rem = i - n * dk * 8, bb = rem >> 3, c = rem & 7, col = 8 * n + c;
    const bool ok = bb < d && col < dr;
    R1s[i] = ok ? E.R1[bb * dr + col] : 0.0;
    R2s[i] = ok ? E.R2[bb * dr + col] : 0.0;
  }
  __syncthreads();
  const double *r1p = R1s + fc * 8 + fr, *r2p = R2s + fc * 8 + fr;
  for (long long item = (long long)blockIdx.x * nw + warp; item < nitems; item += (long long)gridDim.x * nw) {
    const long long mat = item / mtr;
    const int i = (int)(item - mat * mtr);
    const double2 *Tp = reinterpret_cast<const double2 *>(T) + (size_t)item * nt * 64 + lane;
    double cre[NTA][2], cimg[NTA][2];
#pragma unroll
    for (int n = 0; n < NTA; ++n) cre[n][0] = cre[n][1] = cimg[n][0] = cimg[n][1] = 0.0;
    double2 a0 = Tp[0], a1 = Tp[32];
    for (int tt = 0; tt < nt; ++tt) {
      const double2 c0 = a0, c1 = a1;
      if (tt + 1 < nt) { a0 = Tp[(tt + 1) * 64]; a1 = Tp[(tt + 1) * 64 + 32]; }
      const double m0y = -c0.y;
#pragma unroll
      for (int n = 0; n < NTA; ++n) {
        const double r1 = r1p[(n * dk + 4 * tt) * 8], r2 = r2p[(n * dk + 4 * tt) * 8];
        dmma884(cre[n][0], cre[n][1], c0.x, r1);       // (L1 Mqq) R1
        dmma884(cimg[n][0], cimg[n][1], c1.x, r1);     // (L2 Mpq) R1
        dmma884(cre[n][0], cre[n][1], c1.y, r2);       // (L2 Mpp) R2
        dmma884(cimg[n][0], cimg[n][1], m0y, r2);      // -(L1 Mqp) R2
      }
    }
    const int r = 8 * i + fr;
    if (r < dr) {
      double2 *out = cm + (size_t)mat * dr * dr + (size_t)r * dr;
#pragma unroll
      for (int n = 0; n < NTA; ++n) {
        const int col = 8 * n + 2 * fc;
        if (col < dr) out[col] = make_double2(0.5 * cre[n][0], 0.5 * cimg[n][0]);
        if (col + 1 < dr) out[col + 1] = make_double2(0.5 * cre[n][1], 0.5 * cimg[n][1]);
      }
    }
  }
}

static cudaError_t launch_rmult(const EngDev &E, long long nmat, const double *T, double2 *cm, int sm_count, cudaStream_t st) {
  const RmultLayout L = make_rmult_layout(E.d, E.dr);
  const size_t smem = sizeof(double) * (size_t)L.total;
  const long long nitems = nmat * ((E.dr + 7) / 8);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)sm_count * per_sm;
  if (grid > (nitems + 7) / 8) grid = (nitems + 7) / 8;
  if (grid < 1) grid = 1;
#define SC_RMULT_CASE(N)                                                                                          \
  case N: {                                                                                                       \
    cudaError_t ce = cudaFuncSetAttribute(k_rmult<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    if (ce != cudaSuccess) return ce;                                                                             \
    k_rmult<N><<<(int)grid, 256, smem, st>>>(E, nitems, T, cm, L);                                                \
    return cudaGetLastError();                                                                                    \
  }
  switch (L.nta) {
    SC_RMULT_CASE(1) SC_RMULT_CASE(2) SC_RMULT_CASE(3) SC_RMULT_CASE(4) SC_RMULT_CASE(5) SC_RMULT_CASE(6) SC_RMULT_CASE(7)
    SC_RMULT_CASE(8) SC_RMULT_CASE(9) SC_RMULT_CASE(10) SC_RMULT_CASE(11) SC_RMULT_CASE(12)
    default: return cudaErrorInvalidValue;
  }
#undef SC_RMULT_CASE
}

// ------------------------------------------------------------------ (q, p, S) path kernels
// outputs per (step, tl):  qp[(step ntb + tl) 2 d ..] = q, p after the step;  aux[(step ntb + tl) 8 + 6] = S, [+ 7] = T + V of
// the 4th stage point (propagators.py:380)
constexpr int PATH_WARPS = 8;
constexpr int PATH_NE = (SC_MAX_DIM + 31) / 32;         // components per lane

// harmonic molecule (potentials.py:581-593): grad = g0 + H0 (q - pos0), V = e0 + g0.dr + dr.H0.dr / 2 - origin.  H0 in shared
// memory with an odd leading dimension (lane = row: conflict free)
__global__ void __launch_bounds__(32 * PATH_WARPS)
k_path_harmonic(EngDev E, PotDev P, double h, int nsteps, int traj0, int ntb, double *__restrict__ qp, double *__restrict__ aux) {
  extern __shared__ __align__(16) double psm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int d = E.d, ldq = d | 1, dp = (d + 1) & ~1;
  double *H0s = psm, *drv = psm + d * ldq + warp * dp;
  for (int i = t; i < d * d; i += blockDim.x) H0s[(i / d) * ldq + (i % d)] = P.hess0[i];
  __syncthreads();
  const int tl = blockIdx.x * PATH_WARPS + warp;
  if (tl >= ntb) return;
  const int traj = traj0 + tl;
  double *rec = E.rec + (size_t)traj * E.rs;
  double q[PATH_NE], p[PATH_NE], im[PATH_NE], x0[PATH_NE], g0[PATH_NE];
#pragma unroll
  for (int k = 0; k < PATH_NE; ++k) {
    const int a = lane + 32 * k;
    const bool ok = a < d;
    q[k] = ok ? rec[a] : 0.0;
    p[k] = ok ? rec[d + a] : 0.0;
    im[k] = ok ? P.imass[a] : 0.0;
    x0[k] = ok ? P.pos0[a] : 0.0;
    g0[k] = ok ? P.grad0[a] : 0.0;
  }
  double S = rec[2 * d];
  const double vconst = P.e0 - P.origin;
  for (int step = 0; step < nsteps; ++step) {
    double qs[PATH_NE], ps[PATH_NE], accq[PATH_NE], accp[PATH_NE];
    double accS = 0.0, e4 = 0.0;
#pragma unroll
    for (int k = 0; k < PATH_NE; ++k) { qs[k] = q[k]; ps[k] = p[k]; accq[k] = accp[k] = 0.0; }
#pragma unroll 1
    for (int s = 1; s <= 4; ++s) {
      const double cnext = (s == 3) ? h : 0.5 * h;
      const double wgt = (s == 1 || s == 4) ? 1.0 : 2.0;
#pragma unroll
      for (int k = 0; k < PATH_NE; ++k) {
        const int a = lane + 32 * k;
        if (a < d) drv[a] = qs[k] - x0[k];
      }
      __syncwarp();
      double tv = (lane == 0) ? -vconst : 0.0, te = (lane == 0) ? vconst : 0.0;    // contributions to T - V and T + V
#pragma unroll
      for (int k = 0; k < PATH_NE; ++k) {
        const int a = lane + 32 * k;
        if (a < d) {
          const double *hr = H0s + a * ldq;
          double hd = 0.0;
          for (int j = 0; j < d; ++j) hd = fma(hr[j], drv[j], hd);
          const double da = qs[k] - x0[k];
          const double vpart = da * g0[k] + 0.5 * da * hd;
          const double kq = ps[k] * im[k], kp = -(g0[k] + hd);
          const double tk = 0.5 * ps[k] * ps[k] * im[k];
          tv += tk - vpart;
          te += tk + vpart;
          accq[k] += wgt * kq;
          accp[k] += wgt * kp;
          if (s < 4) {
            qs[k] = q[k] + cnext * kq;
            ps[k] = p[k] + cnext * kp;
          }
        }
      }
      accS += wgt * tv;
      if (s == 4) e4 = te;
      __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      accS += __shfl_xor_sync(0xffffffffu, accS, o);
      e4 += __shfl_xor_sync(0xffffffffu, e4, o);
    }
    S += h / 6.0 * accS;
    double *qo = qp + ((size_t)step * ntb + tl) * 2 * d;
#pragma unroll
    for (int k = 0; k < PATH_NE; ++k) {
      const int a = lane + 32 * k;
      q[k] += h / 6.0 * accq[k];
      p[k] += h / 6.0 * accp[k];
      if (a < d) { qo[a] = q[k]; qo[d + a] = p[k]; }
    }
    if (lane == 0) {
      double *ax = aux + ((size_t)step * ntb + tl) * 8;
      ax[6] = S;
      ax[7] = e4;
    }
  }
#pragma unroll
  for (int k = 0; k < PATH_NE; ++k) {
    const int a = lane + 32 * k;
    if (a < d) { rec[a] = q[k]; rec[d + a] = p[k]; }
  }
  if (lane == 0) rec[2 * d] = S;
}

// separable potentials (Morse / AS, NonHarmonic) with ANY width matrices: like k_qp_path (sc_chunk.cuh) without its
// diagonal-width overlap terms (k_aux_terms computes them from the stored q, p).  hd (step, tl, 4, dp): stage Hessian diagonals.
__global__ void __launch_bounds__(32 * PATH_WARPS)
k_path_separable(EngDev E, PotDev P, double h, int nsteps, int traj0, int ntb, double *__restrict__ qp, double *__restrict__ aux,
                 double *__restrict__ hd) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = E.d, dp = (d + 1) & ~1;
  const int tl = blockIdx.x * PATH_WARPS + warp;
  if (tl >= ntb) return;
  double *rec = E.rec + (size_t)(traj0 + tl) * E.rs;
  double q[PATH_NE], p[PATH_NE], im[PATH_NE];
#pragma unroll
  for (int k = 0; k < PATH_NE; ++k) {
    const int a = lane + 32 * k;
    const bool ok = a < d;
    q[k] = ok ? rec[a] : 0.0;
    p[k] = ok ? rec[d + a] : 0.0;
    im[k] = ok ? P.imass[a] : 0.0;
  }
  double S = rec[2 * d];
  for (int step = 0; step < nsteps; ++step) {
    double *hrow = hd + ((size_t)step * ntb + tl) * 4 * dp;
    double accS = 0.0, e4 = 0.0;
#pragma unroll
    for (int k = 0; k < PATH_NE; ++k) {
      const int a = lane + 32 * k;
      if (a < d) {
        const double qa = q[k], pa = p[k];
        double qsa = qa, psa = pa, accq = 0.0, accp = 0.0;
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
          if (s == 4) e4 += tk + vpart;
          accq += wgt * kq;
          accp += wgt * kp;
          if (s < 4) {
            qsa = qa + cnext * kq;
            psa = pa + cnext * kp;
          }
        }
        q[k] = qa + h / 6.0 * accq;
        p[k] = pa + h / 6.0 * accp;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      accS += __shfl_xor_sync(0xffffffffu, accS, o);
      e4 += __shfl_xor_sync(0xffffffffu, e4, o);
    }
    S += h / 6.0 * accS;
    double *qo = qp + ((size_t)step * ntb + tl) * 2 * d;
#pragma unroll
    for (int k = 0; k < PATH_NE; ++k) {
      const int a = lane + 32 * k;
      if (a < d) { qo[a] = q[k]; qo[d + a] = p[k]; }
    }
    if (lane == 0) {
      double *ax = aux + ((size_t)step * ntb + tl) * 8;
      ax[6] = S;
      ax[7] = e4;
    }
  }
#pragma unroll
  for (int k = 0; k < PATH_NE; ++k) {
    const int a = lane + 32 * k;
    if (a < d) { rec[a] = q[k]; rec[d + a] = p[k]; }
  }
  if (lane == 0) rec[2 * d] = S;
}

// rotated Morse potential (dense parity fixture): V'(x) = V(Q^T x); r = Q^T x, inner Morse per mode,
// grad = Q g.  One warp per trajectory; the inner second derivatives h_k of every stage go to hd (step, tl, 4, dp) --
// k_expand_hessian forms H_s = Q diag(h_s) Q^T on the tensor pipe.  Q in shared memory with an odd leading dimension
// (column access: consecutive lanes; row access: stride ldq).
__global__ void __launch_bounds__(32 * PATH_WARPS)
k_path_rotated(EngDev E, PotDev P, double h, int nsteps, int traj0, int ntb, double *__restrict__ qp, double *__restrict__ aux,
               double *__restrict__ hd) {
  extern __shared__ __align__(16) double psm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int d = E.d, ldq = d | 1, dp = (d + 1) & ~1;
  double *Qs = psm, *xv = psm + d * ldq + warp * 2 * dp, *gv = xv + dp;
  for (int i = t; i < d * d; i += blockDim.x) Qs[(i / d) * ldq + (i % d)] = P.Q[i];
  __syncthreads();
  const int tl = blockIdx.x * PATH_WARPS + warp;
  if (tl >= ntb) return;
  const int traj = traj0 + tl;
  double *rec = E.rec + (size_t)traj * E.rs;
  double q[PATH_NE], p[PATH_NE], im[PATH_NE], pa_[PATH_NE], pD[PATH_NE], pw2[PATH_NE];
#pragma unroll
  for (int k = 0; k < PATH_NE; ++k) {
    const int a = lane + 32 * k;
    const bool ok = a < d;
    q[k] = ok ? rec[a] : 0.0;
    p[k] = ok ? rec[d + a] : 0.0;
    im[k] = ok ? P.imass[a] : 0.0;
    pa_[k] = (ok && !P.all_harmonic) ? P.a[a] : 0.0;
    pD[k] = (ok && !P.all_harmonic) ? P.D[a] : 0.0;
    pw2[k] = ok ? P.omega[a] * P.omega[a] : 0.0;
  }
  double S = rec[2 * d];
  for (int step = 0; step < nsteps; ++step) {
    double qs[PATH_NE], ps[PATH_NE], accq[PATH_NE], accp[PATH_NE];
    double accS = 0.0, e4 = 0.0;
#pragma unroll
    for (int k = 0; k < PATH_NE; ++k) { qs[k] = q[k]; ps[k] = p[k]; accq[k] = accp[k] = 0.0; }
    double *hrow = hd + ((size_t)step * ntb + tl) * 4 * dp;
#pragma unroll 1
    for (int s = 1; s <= 4; ++s) {
      const double cnext = (s == 3) ? h : 0.5 * h;
      const double wgt = (s == 1 || s == 4) ? 1.0 : 2.0;
#pragma unroll
      for (int k = 0; k < PATH_NE; ++k) {
        const int a = lane + 32 * k;
        if (a < d) xv[a] = qs[k];
      }
      __syncwarp();
      double vsum = (lane == 0) ? -P.origin : 0.0;
#pragma unroll
      for (int k = 0; k < PATH_NE; ++k) {
        const int m = lane + 32 * k;
        if (m < d) {
          double r = 0.0;
          for (int i = 0; i < d; ++i) r = fma(Qs[i * ldq + m], xv[i], r);
          double gi, hi;
          if (P.all_harmonic) {
            vsum += 0.5 * pw2[k] * r * r;
            gi = pw2[k] * r;
            hi = pw2[k];
          } else {
            const double e = exp(-pa_[k] * r);
            vsum += pD[k] * (1.0 - e) * (1.0 - e);
            gi = 2.0 * pa_[k] * pD[k] * e * (1.0 - e);
            hi = 2.0 * pa_[k] * pa_[k] * pD[k] * e * (2.0 * e - 1.0);
          }
          gv[m] = gi;
          hrow[(s - 1) * dp + m] = hi;
        }
      }
      __syncwarp();
      double tv = -vsum, te = vsum;
#pragma unroll
      for (int k = 0; k < PATH_NE; ++k) {
        const int a = lane + 32 * k;
        if (a < d) {
          const double *qr = Qs + a * ldq;
          double g = 0.0;
          for (int m = 0; m < d; ++m) g = fma(qr[m], gv[m], g);
          const double kq = ps[k] * im[k], kp = -g;
          const double tk = 0.5 * ps[k] * ps[k] * im[k];
          tv += tk;
          te += tk;
          accq[k] += wgt * kq;
          accp[k] += wgt * kp;
          if (s < 4) {
            qs[k] = q[k] + cnext * kq;
            ps[k] = p[k] + cnext * kp;
          }
        }
      }
      accS += wgt * tv;
      if (s == 4) e4 = te;
      __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      accS += __shfl_xor_sync(0xffffffffu, accS, o);
      e4 += __shfl_xor_sync(0xffffffffu, e4, o);
    }
    S += h / 6.0 * accS;
    double *qo = qp + ((size_t)step * ntb + tl) * 2 * d;
#pragma unroll
    for (int k = 0; k < PATH_NE; ++k) {
      const int a = lane + 32 * k;
      q[k] += h / 6.0 * accq[k];
      p[k] += h / 6.0 * accp[k];
      if (a < d) { qo[a] = q[k]; qo[d + a] = p[k]; }
    }
    if (lane == 0) {
      double *ax = aux + ((size_t)step * ntb + tl) * 8;
      ax[6] = S;
      ax[7] = e4;
    }
  }
#pragma unroll
  for (int k = 0; k < PATH_NE; ++k) {
    const int a = lane + 32 * k;
    if (a < d) { rec[a] = q[k]; rec[d + a] = p[k]; }
  }
  if (lane == 0) rec[2 * d] = S;
}

// (q, p, S) path of potentials evaluated by their own batched kernel (sGDML, sc_gdml.cuh): one thread per trajectory, state
// in batch-last arrays (coalesced over the trajectories) of the path scratch pst = [qa | pa | qs | ps | accq | accp] (each d x nt)
// + [S | accS] (nt).  k_gstage_begin loads the records and emits the stage-1 positions; k_gstage_adv consumes V, grad of stage s
// (classical RK4, propagators.py:114-119, 361-368) and emits the next positions, after stage 4 the per-step outputs.
__global__ void k_gstage_begin(EngDev E, int traj0, int nt, double *__restrict__ pst, double *__restrict__ r_out) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int d = E.d;
  const double *rec = E.rec + (size_t)(traj0 + tl) * E.rs;
  const size_t dn = (size_t)d * nt;
  for (int a = 0; a < d; ++a) {
    const double q = rec[a], p = rec[d + a];
    const size_t i = (size_t)a * nt + tl;
    pst[i] = q; pst[dn + i] = p; pst[2 * dn + i] = q; pst[3 * dn + i] = p; pst[4 * dn + i] = 0.0; pst[5 * dn + i] = 0.0;
    r_out[i] = q;
  }
  pst[6 * dn + tl] = rec[2 * d];
  pst[6 * dn + nt + tl] = 0.0;
}

__global__ void k_gstage_adv(EngDev E, PotDev P, double h, int s, int step, int last_step, int traj0, int nt, double *__restrict__ pst,
                             const double *__restrict__ V, const double *__restrict__ grad, double *__restrict__ r_out,
                             double *__restrict__ qp, double *__restrict__ aux) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int d = E.d;
  const size_t dn = (size_t)d * nt;
  const double cnext = (s == 3) ? h : 0.5 * h;
  const double wgt = (s == 1 || s == 4) ? 1.0 : 2.0;
  double tk = 0.0;
  double *qo = qp + ((size_t)step * nt + tl) * 2 * d;
  double *rec = E.rec + (size_t)(traj0 + tl) * E.rs;
  for (int a = 0; a < d; ++a) {
    const size_t i = (size_t)a * nt + tl;
    const double im = P.imass[a], ps = pst[3 * dn + i];
    const double kq = ps * im, kp = -grad[i];
    tk += 0.5 * ps * ps * im;
    const double aq = pst[4 * dn + i] + wgt * kq, ap = pst[5 * dn + i] + wgt * kp;
    if (s < 4) {
      pst[4 * dn + i] = aq;
      pst[5 * dn + i] = ap;
      const double qs = pst[i] + cnext * kq;
      pst[2 * dn + i] = qs;
      pst[3 * dn + i] = pst[dn + i] + cnext * kp;
      r_out[i] = qs;
    } else {
      const double q = pst[i] + h / 6.0 * aq, p = pst[dn + i] + h / 6.0 * ap;
      pst[i] = q; pst[dn + i] = p; pst[2 * dn + i] = q; pst[3 * dn + i] = p; pst[4 * dn + i] = 0.0; pst[5 * dn + i] = 0.0;
      r_out[i] = q;
      qo[a] = q;
      qo[d + a] = p;
      if (last_step) { rec[a] = q; rec[d + a] = p; }
    }
  }
  const double v = V[tl];
  const double aS = pst[6 * dn + nt + tl] + wgt * (tk - v);
  if (s < 4) {
    pst[6 * dn + nt + tl] = aS;
  } else {
    const double S = pst[6 * dn + tl] + h / 6.0 * aS;
    pst[6 * dn + tl] = S;
    pst[6 * dn + nt + tl] = 0.0;
    double *ax = aux + ((size_t)step * nt + tl) * 8;
    ax[6] = S;
    ax[7] = tk + v;
    if (last_step) rec[2 * d] = S;
  }
}

// Hessian stream images (d x LDH, zero padded) from the stage diagonals hd (item, dp), item = (step, tl, stage):
//   ROT = 0  H = diag(h)                    (separable models run through the general dense engine)
//   ROT = 1  H = Q diag(h) Q^T              (rotated Morse fixture) on the tensor pipe: A fragment = Q[i][k] h[k], B fragment =
//            Q[j][k], upper-triangular 8 x 8 tiles + mirrored stores.  Q in shared memory, ld = LDH (conflict-free fragments)
// The image of item (step, tl, s) goes to hs + ((step 4 + s) ntb + tl) hsz.
template <int NK, int ROT>
__global__ void __launch_bounds__(256)
k_expand_hessian(PotDev P, int nsteps, int ntb, const double *__restrict__ hd, double *__restrict__ hs) {
  constexpr int MT = (NK + 1) / 2, LDH = cols_ldh(NK);
  extern __shared__ __align__(16) double esm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int d = P.d, dp = (d + 1) & ~1, hsz = d * LDH;
  const int fr = lane >> 2, fc = lane & 3;
  double *Qs = esm, *hv = esm + (ROT ? 8 * MT * LDH : 0) + warp * 4 * NK;
  if (ROT) {
    for (int i = t; i < 8 * MT * LDH; i += blockDim.x) {
      const int r = i / LDH, c = i - r * LDH;
      Qs[i] = (r < d && c < d) ? P.Q[r * d + c] : 0.0;
    }
    __syncthreads();
  }
  const long long nitems = (long long)nsteps * ntb * 4;
  for (long long item = (long long)blockIdx.x * 8 + warp; item < nitems; item += (long long)gridDim.x * 8) {
    const long long st_tl = item >> 2;
    const int s = (int)(item & 3), step = (int)(st_tl / ntb), tl = (int)(st_tl - (long long)step * ntb);
    const double *hrow = hd + (size_t)item * dp;
    double *H = hs + ((size_t)(step * 4 + s) * ntb + tl) * hsz;
    if (!ROT) {
      for (int i = lane; i < hsz; i += 32) {
        const int r = i / LDH, c = i - r * LDH;
        H[i] = (r == c) ? hrow[r] : 0.0;
      }
      continue;
    }
    __syncwarp();
    for (int k = lane; k < 4 * NK; k += 32) hv[k] = k < d ? hrow[k] : 0.0;
    __syncwarp();
    // zero padding columns d .. LDH - 1
    for (int i = lane; i < d * (LDH - d); i += 32) {
      const int r = i / (LDH - d), c = d + i - r * (LDH - d);
      H[r * LDH + c] = 0.0;
    }
#pragma unroll 1
    for (int I = 0; I < MT; ++I) {
      double a[NK];
#pragma unroll
      for (int kk = 0; kk < NK; ++kk) a[kk] = Qs[(8 * I + fr) * LDH + 4 * kk + fc] * hv[4 * kk + fc];
#pragma unroll 1
      for (int J = I; J < MT; ++J) {
        double c0 = 0.0, c1 = 0.0;
        const double *bq = Qs + (8 * J + fr) * LDH + fc;
#pragma unroll
        for (int kk = 0; kk < NK; ++kk) dmma884(c0, c1, a[kk], bq[4 * kk]);
        const int row = 8 * I + fr, col = 8 * J + 2 * fc;
        if (row < d) {
          if (col < d) H[row * LDH + col] = c0;
          if (col + 1 < d) H[row * LDH + col + 1] = c1;
        }
        if (J > I && row < d) {
          if (col < d) H[col * LDH + row] = c0;
          if (col + 1 < d) H[(col + 1) * LDH + row] = c1;
        }
      }
    }
  }
}

template <int NK>
static cudaError_t launch_expand_t(const PotDev &P, int rot, int nsteps, int ntb, const double *hd, double *hs, int sm_count,
                                   cudaStream_t st) {
  constexpr int MT = (NK + 1) / 2, LDH = cols_ldh(NK);
  const long long nitems = (long long)nsteps * ntb * 4;
  long long grid = std::min<long long>((nitems + 7) / 8, (long long)sm_count * (rot ? 2 : 8));
  if (grid < 1) grid = 1;
  if (rot) {
    const size_t smem = sizeof(double) * ((size_t)8 * MT * LDH + 8 * 4 * NK);
    cudaError_t ce = cudaFuncSetAttribute(k_expand_hessian<NK, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return ce;
    k_expand_hessian<NK, 1><<<(int)grid, 256, smem, st>>>(P, nsteps, ntb, hd, hs);
  } else {
    k_expand_hessian<NK, 0><<<(int)grid, 256, 0, st>>>(P, nsteps, ntb, hd, hs);
  }
  return cudaGetLastError();
}

static cudaError_t launch_expand(const PotDev &P, int rot, int nsteps, int ntb, const double *hd, double *hs, int sm_count,
                                 cudaStream_t st) {
  switch ((P.d + 3) / 4) {
#define SC_EXPAND_CASE(N) case N: return launch_expand_t<N>(P, rot, nsteps, ntb, hd, hs, sm_count, st);
    SC_EXPAND_CASE(4) SC_EXPAND_CASE(5) SC_EXPAND_CASE(6) SC_EXPAND_CASE(7) SC_EXPAND_CASE(8) SC_EXPAND_CASE(9) SC_EXPAND_CASE(10)
    SC_EXPAND_CASE(11) SC_EXPAND_CASE(12) SC_EXPAND_CASE(13) SC_EXPAND_CASE(14) SC_EXPAND_CASE(15) SC_EXPAND_CASE(16)
#undef SC_EXPAND_CASE
    default: return cudaErrorInvalidValue;
  }
}

// overlap / NAC partial sums v[0..5] (corr_terms, sc_device.cuh) of every (step, trajectory) from the stored q, p; one warp
// per item, dense or diagonal overlap matrices
__global__ void __launch_bounds__(256)
k_aux_terms(EngDev E, int nsteps, int traj0, int ntb, const double *__restrict__ qp, double *__restrict__ aux) {
  __shared__ double sh[8][2 * SC_MAX_DIM];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, d = E.d;
  const long long nitems = (long long)nsteps * ntb;
  double *dqv = sh[warp], *dpv = sh[warp] + d;
  for (long long item = (long long)blockIdx.x * 8 + warp; item < nitems; item += (long long)gridDim.x * 8) {
    const int tl = (int)(item % ntb);
    const double *q = qp + (size_t)item * 2 * d, *p = q + d;
    const double *zt = E.zt + (size_t)(traj0 + tl) * 2 * d;
    __syncwarp();
    for (int a = lane; a < d; a += 32) { dqv[a] = E.q0[a] - q[a]; dpv[a] = E.p0[a] - p[a]; }
    __syncwarp();
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int a = lane; a < d; a += 32) {
      const double dq = dqv[a], dpa = dpv[a];
      if (E.diag) {
        v[0] += -0.5 * (dq * E.otA[a] * dq + dpa * E.otB[a] * dpa);
        v[1] += -E.p0[a] * dq + dq * E.otC[a] * dpa;
      } else {
        double sa = 0.0, sb = 0.0, sc_ = 0.0;
        for (int j = 0; j < d; ++j) {
          sa = fma(__ldg(E.otA + j * d + a), dqv[j], sa);
          sb = fma(__ldg(E.otB + j * d + a), dpv[j], sb);
          sc_ = fma(__ldg(E.otC + j * d + a), dqv[j], sc_);
        }
        v[0] += -0.5 * (dq * sa + dpa * sb);
        v[1] += -E.p0[a] * dq + dpa * sc_;
      }
      const double wr = E.wR[a], wg = E.wG[a];
      v[2] += dq * wr;
      v[3] += -dpa * wg;
      v[4] += (E.q0[a] - zt[a]) * wr;
      v[5] += (zt[d + a] - E.p0[a]) * wg;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
    if (lane < 6) {
      double x = v[0];
#pragma unroll
      for (int i = 1; i < 6; ++i) x = (lane == i) ? v[i] : x;
      aux[(size_t)item * 8 + lane] = x;
    }
  }
}

// branch tracking without propagation (MODE_INIT / MODE_TRACK of sc_kernels.cuh for d the generic kernel cannot hold)
__global__ void k_track_only(EngDev E, int traj0, int nt, const double2 *__restrict__ det, int init) {
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  if (tl >= nt) return;
  const int traj = traj0 + tl;
  const double2 dtv = det[tl];
  E.sign[traj] = init ? 1.0 : track_sign(E.sign[traj], E.c2[traj], dtv);
  E.c2[traj] = dtv;
  E.c[traj] = csqrt_principal(dtv);
}

// contributions of the CURRENT state to C_auto and k_ic (MODE_CORR): one warp per trajectory, per-block partial rows
__global__ void __launch_bounds__(256)
k_corr_now(EngDev E, double *__restrict__ partials) {
  __shared__ double sh[8][2 * SC_MAX_DIM];
  __shared__ double red[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, d = E.d;
  double *dqv = sh[warp], *dpv = sh[warp] + d;
  double acc4[4] = {0.0, 0.0, 0.0, 0.0};
  for (int traj = blockIdx.x * 8 + warp; traj < E.n; traj += gridDim.x * 8) {
    const double *rec = E.rec + (size_t)traj * E.rs;
    const double *zt = E.zt + (size_t)traj * 2 * d;
    __syncwarp();
    for (int a = lane; a < d; a += 32) { dqv[a] = E.q0[a] - rec[a]; dpv[a] = E.p0[a] - rec[d + a]; }
    __syncwarp();
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int a = lane; a < d; a += 32) {
      const double dq = dqv[a], dpa = dpv[a];
      if (E.diag) {
        v[0] += -0.5 * (dq * E.otA[a] * dq + dpa * E.otB[a] * dpa);
        v[1] += -E.p0[a] * dq + dq * E.otC[a] * dpa;
      } else {
        double sa = 0.0, sb = 0.0, sc_ = 0.0;
        for (int j = 0; j < d; ++j) {
          sa = fma(__ldg(E.otA + j * d + a), dqv[j], sa);
          sb = fma(__ldg(E.otB + j * d + a), dpv[j], sb);
          sc_ = fma(__ldg(E.otC + j * d + a), dqv[j], sc_);
        }
        v[0] += -0.5 * (dq * sa + dpa * sb);
        v[1] += -E.p0[a] * dq + dpa * sc_;
      }
      const double wr = E.wR[a], wg = E.wG[a];
      v[2] += dq * wr;
      v[3] += -dpa * wg;
      v[4] += (E.q0[a] - zt[a]) * wr;
      v[5] += (zt[d + a] - E.p0[a]) * wg;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
    if (lane == 0) {
      double2 ca, ki;
      corr_finish(E, v, rec[2 * d], E.c[traj], E.sign[traj], E.wvi[traj], ca, ki);
      acc4[0] += ca.x; acc4[1] += ca.y; acc4[2] += ki.x; acc4[3] += ki.y;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red[warp][i] = acc4[i];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0.0;
    if (threadIdx.x < 4)
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 5 + threadIdx.x] = s;
  }
}

// separable models too: their diagonal stage Hessians are expanded to full matrices (the kernel multiplies whatever it is
// given).  With diagonal full-rank widths and d > 32 the structured pipeline of sc_chunk.cuh is faster and is preferred by
// the dispatcher unless the dense engine is asked for (option dense_engine)
// contributions of the CURRENT state with POSITION-DEPENDENT non-adiabatic couplings (propagators.py:868-909 in full
// generality): n1Q, n1q (d, n) = -hbar^2 tau1 / m at the current / initial positions, n2Q, n2q (n) = -hbar^2/2 sum_k tau2_k / m_k;
//   nacQ = n2Q + (q0 - Q) R n1Q - i PI . n1Q,   PI = p0 + G0 iGi0 (P - p0)
//   nacq = n2q + (q0 - q) R n1q + i pi . n1q,   pi = p0 + G0 iGi0 (p - p0),      R = G0 iGi0 Gi
// One warp per trajectory, per-block partial rows (C_auto re, im, k_ic re, im, 0).
__global__ void __launch_bounds__(256)
k_corr_general(EngDev E, const double *__restrict__ R, const double *__restrict__ G0iG, const double *__restrict__ n1Q,
               const double *__restrict__ n1q, const double *__restrict__ n2Q, const double *__restrict__ n2q,
               double *__restrict__ partials) {
  __shared__ double sh[8][6 * SC_MAX_DIM];
  __shared__ double red[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, d = E.d, n = E.n;
  double *dQv = sh[warp], *dPv = dQv + d, *dqv = dPv + d, *dpv = dqv + d, *nQ = dpv + d, *nq = nQ + d;
  double acc4[4] = {0.0, 0.0, 0.0, 0.0};
  for (int traj = blockIdx.x * 8 + warp; traj < n; traj += gridDim.x * 8) {
    const double *rec = E.rec + (size_t)traj * E.rs;
    const double *zt = E.zt + (size_t)traj * 2 * d;
    __syncwarp();
    for (int a = lane; a < d; a += 32) {
      dQv[a] = E.q0[a] - rec[a];            // q0 - Q
      dPv[a] = E.p0[a] - rec[d + a];        // p0 - P
      dqv[a] = E.q0[a] - zt[a];             // q0 - q
      dpv[a] = E.p0[a] - zt[d + a];         // p0 - p
      nQ[a] = n1Q[(size_t)a * n + traj];
      nq[a] = n1q[(size_t)a * n + traj];
    }
    __syncwarp();
    double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int a = lane; a < d; a += 32) {
      const double dq = dQv[a], dpa = dPv[a];
      if (E.diag) {
        v[0] += -0.5 * (dq * E.otA[a] * dq + dpa * E.otB[a] * dpa);
        v[1] += -E.p0[a] * dq + dq * E.otC[a] * dpa;
      } else {
        double sa = 0.0, sb = 0.0, sc_ = 0.0;
        for (int j = 0; j < d; ++j) {
          sa = fma(__ldg(E.otA + j * d + a), dQv[j], sa);
          sb = fma(__ldg(E.otB + j * d + a), dPv[j], sb);
          sc_ = fma(__ldg(E.otC + j * d + a), dQv[j], sc_);
        }
        v[0] += -0.5 * (dq * sa + dpa * sb);
        v[1] += -E.p0[a] * dq + dpa * sc_;
      }
      // row a of R n1 and of G0 iGi0 (p0 - P): coalesced over the lanes through the transposed access (column a)
      double rQ = 0.0, rq = 0.0, gP = 0.0, gp = 0.0;
      for (int j = 0; j < d; ++j) {
        const double r = __ldg(R + (size_t)a * d + j), g = __ldg(G0iG + (size_t)a * d + j);
        rQ = fma(r, nQ[j], rQ);
        rq = fma(r, nq[j], rq);
        gP = fma(g, dPv[j], gP);
        gp = fma(g, dpv[j], gp);
      }
      v[2] += dq * rQ;                                   // (q0 - Q) R n1Q
      v[3] += (E.p0[a] - gP) * nQ[a];                    // PI . n1Q,  PI_a = p0_a - [G0 iGi0 (p0 - P)]_a
      v[4] += dqv[a] * rq;                               // (q0 - q) R n1q
      v[5] += (E.p0[a] - gp) * nq[a];                    // pi . n1q
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
    if (lane == 0) {
      const double2 c = E.c[traj], wvi = E.wvi[traj];
      const double sg = E.sign[traj];
      const double2 e = cexp(v[0], rec[2 * d] - v[1]);
      double2 cq = cmul(make_double2(E.ot_fac * e.x, E.ot_fac * e.y), wvi);
      cq = cmul(cq, make_double2(sg * c.x, sg * c.y));
      const double2 nacQ = make_double2(n2Q[traj] + v[2], -v[3]);
      const double2 nacq = make_double2(n2q[traj] + v[4], v[5]);
      const double2 k = cmul(cmul(nacQ, nacq), cq);
      acc4[0] += cq.x; acc4[1] += cq.y; acc4[2] += k.x; acc4[3] += k.y;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red[warp][i] = acc4[i];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0.0;
    if (threadIdx.x < 4)
      for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 5 + threadIdx.x] = s;
  }
}

static bool stream_supported(const EngDev &E, const PotDev &P) {
  if (E.d < 13 || E.d > SC_MAX_DIM) return false;      // 13 .. 16: between the largest k_hk_small and the stage interface's limit
  if (P.type == POT_GDML && E.d < 17) return false;    // the sGDML stream starts where the stage interface switches (d >= 17)
  if (P.type == POT_HARMONIC) return true;
  if (E.d > 64) return false;                  // Hessian expansion / sGDML kernels: d <= 64
  return P.type == POT_ROTATED_MORSE || P.type == POT_GDML || P.type == POT_MORSE || P.type == POT_NONHARMONIC;
}

}  // namespace sc
