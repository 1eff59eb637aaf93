// sc_stage.cuh -- generic-potential stage interface: any object implementing the reference's potential protocol
// (potentials.py: harmonic_approximation(r) -> V (n), grad (d, n), hess (d, d, n)) drives one classical RK4 step
// (propagators.py:86-119, 313-383) as
//     4 x { sc_engine_stage_positions -> caller evaluates V, grad, hess -> sc_engine_stage_apply },  sc_engine_stage_finish
// The stage state ys and the weighted sum of the stage derivatives live in two ensemble-sized global buffers in the
// record layout; the prefactor / branch tracking of the new time is done by the fused kernel (MODE_TRACK).
// Included by sc_engine.cu after the engine definitions.

namespace sc {

// q_out (d, n) batch-last <- positions of the stage state
__global__ void k_stage_positions(EngDev E, const double *src, double *q_out) {
  const int d = E.d, n = E.n;
  const size_t total = (size_t)d * n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int a = (int)(i / n), traj = (int)(i % n);
    q_out[i] = src[(size_t)traj * E.rs + a];
  }
}

// one CTA per trajectory (grid-stride): k_s = f(ys_s); acc += w_s k_s; ys_{s+1} = y + c_{s+1} k_s or y += h/6 acc
__global__ void __launch_bounds__(256)
k_stage_apply(EngDev E, int stage, double h, const double *masses, const double *V, const double *grad, const double *hess,
              double *ys, double *acc, double *esum) {
  extern __shared__ __align__(16) double st_smem[];
  const int d = E.d, n = E.n, W = 2 * d, NE = 2 * d * d, t = threadIdx.x, NT = blockDim.x;
  double *Us = st_smem, *Vs = Us + NE, *H = Vs + NE, *ps = H + d * d, *im = ps + d, *red = im + d;
  const double w = (stage == 1 || stage == 4) ? 1.0 : 2.0;
  const double cn = (stage == 3) ? h : 0.5 * h;
  for (int i = t; i < d; i += NT) im[i] = 1.0 / masses[i];
  for (int traj = blockIdx.x; traj < n; traj += gridDim.x) {
    double *rec = E.rec + (size_t)traj * E.rs;
    double *y_s = ys + (size_t)traj * E.rs;
    double *ac = acc + (size_t)traj * E.rs;
    const double *src = (stage == 1) ? rec : y_s;
    __syncthreads();
    for (int idx = t; idx < NE; idx += NT) { Us[idx] = src[E.qps + idx]; Vs[idx] = src[E.qps + NE + idx]; }
    for (int idx = t; idx < d * d; idx += NT) H[idx] = hess[(size_t)idx * n + traj];
    for (int i = t; i < d; i += NT) ps[i] = src[d + i];
    __syncthreads();
    for (int idx = t; idx < NE; idx += NT) {
      const int a = idx / W, b = idx % W;
      const double kU = Vs[idx] * im[a];
      double s = 0.0;
      for (int k = 0; k < d; ++k) s = fma(H[a * d + k], Us[k * W + b], s);
      const double kV = -s;
      const double aU = (stage == 1 ? 0.0 : ac[E.qps + idx]) + w * kU;
      const double aV = (stage == 1 ? 0.0 : ac[E.qps + NE + idx]) + w * kV;
      if (stage < 4) {
        ac[E.qps + idx] = aU;
        ac[E.qps + NE + idx] = aV;
        y_s[E.qps + idx] = rec[E.qps + idx] + cn * kU;
        y_s[E.qps + NE + idx] = rec[E.qps + NE + idx] + cn * kV;
      } else {
        rec[E.qps + idx] += h / 6.0 * aU;
        rec[E.qps + NE + idx] += h / 6.0 * aV;
      }
    }
    // vector part: q, p, S
    double tk = 0.0;
    for (int a = t; a < d; a += NT) {
      const double kq = ps[a] * im[a], kp = -grad[(size_t)a * n + traj];
      tk += 0.5 * ps[a] * ps[a] * im[a];
      const double aq = (stage == 1 ? 0.0 : ac[a]) + w * kq;
      const double ap = (stage == 1 ? 0.0 : ac[d + a]) + w * kp;
      if (stage < 4) {
        ac[a] = aq;
        ac[d + a] = ap;
        y_s[a] = rec[a] + cn * kq;
        y_s[d + a] = rec[d + a] + cn * kp;
      } else {
        rec[a] += h / 6.0 * aq;
        rec[d + a] += h / 6.0 * ap;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tk += __shfl_xor_sync(0xffffffffu, tk, o);
    if ((t & 31) == 0) red[t >> 5] = tk;
    __syncthreads();
    if (t == 0) {
      double T = 0.0;
      for (int k = 0; k < (NT >> 5); ++k) T += red[k];
      const double v = V[traj];
      const double aS = (stage == 1 ? 0.0 : ac[2 * d]) + w * (T - v);
      if (stage < 4) {
        ac[2 * d] = aS;
      } else {
        rec[2 * d] += h / 6.0 * aS;
        if (esum) atomicAdd(esum, T + v);     // <T+V> of the 4th stage point (propagators.py:380)
      }
    }
  }
}

}  // namespace sc

namespace sc {

// hess (d, d, n) batch-last (the potential protocol's layout, potentials.py: harmonic_approximation) -> the stream images
// (d x ldh, zero padded) k_rk4_stream consumes; image of trajectory tl at out + tl hsz.  One warp per trajectory row.
__global__ void __launch_bounds__(256)
k_hess_to_image(int d, int n, int ldh, const double *__restrict__ hess, double *__restrict__ out) {
  const size_t hsz = (size_t)d * ldh;
  const size_t total = (size_t)n * hsz;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // consecutive threads: consecutive trajectories of one matrix element (coalesced reads); the writes are strided by hsz
    const size_t e = i / n;
    const int tl = (int)(i - e * n), r = (int)(e / ldh), c = (int)(e - (size_t)r * ldh);
    out[(size_t)tl * hsz + e] = c < d ? hess[((size_t)r * d + c) * n + tl] : 0.0;
  }
}

__global__ void k_inv_masses(int d, const double *__restrict__ m, double *__restrict__ im) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d) im[i] = 1.0 / m[i];
}

// <T + V> of the 4th stage point, summed over the ensemble (propagators.py:380): aux rows written by k_gstage_adv
__global__ void k_stage_energy(int n, const double *__restrict__ aux, double *__restrict__ esum) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += aux[(size_t)i * 8 + 7];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(esum, t);
  }
}

}  // namespace sc

// 17 <= d <= SC_MAX_DIM: the stage interface runs on the dense column pipeline -- the caller's Hessians become the stream
// images of the four RK4 stages, the (q, p, S) path is advanced by k_gstage_adv, and sc_engine_stage_finish propagates the
// monodromy blocks with k_rk4_stream (one step) and takes the prefactor through k_rmult / the batched LU / k_track_only.
// Buffers (ensemble sized, kept for the next step): path state | positions | images of 4 stages | q, p | aux | 1/m
static bool stage_uses_stream(const sc_engine *e) { return e->dev.d >= 17 && e->dev.d <= SC_MAX_DIM && !getenv("SC_NO_STREAM"); }

static int stage_stream_buffers(sc_engine *e, StreamLayout &L, cudaStream_t st) {
  if (int rc = stream_setup(e, L, st)) return rc;
  if (e->stage_buf) return SC_OK;
  const size_t n = e->dev.n, d = e->dev.d;
  CU(e->ens.alloc(n * (6 * d + 2) + n * 4 * (size_t)L.hsz + n * 2 * d + n * 8 + d + 16, &e->stage_buf));
  return SC_OK;
}

static int stage_buffers(sc_engine *e) {
  if (e->stage_buf) return SC_OK;
  CU(e->ens.alloc((size_t)2 * e->dev.n * e->dev.rs, &e->stage_buf));
  return SC_OK;
}

extern "C" int sc_engine_stage_positions(sc_engine *e, int stage, double dt, double *q_dev, void *stream) {
  (void)dt;
  if (!e || !q_dev || e->dev.n < 1) return fail(SC_ERR_INVALID, "no ensemble / null argument");
  if (stage < 1 || stage > 4) return fail(SC_ERR_INVALID, "stage %d outside 1..4", stage);
  if (stage_uses_stream(e)) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamLayout L;
    if (int rc = stage_stream_buffers(e, L, st)) return rc;
    const size_t n = e->dev.n, d = e->dev.d;
    double *pst = e->stage_buf;
    if (stage == 1) {
      k_gstage_begin<<<(int)((n + 127) / 128), 128, 0, st>>>(e->dev, 0, (int)n, pst, q_dev);
      CU(cudaGetLastError());
    } else {
      CU(cudaMemcpyAsync(q_dev, pst + 2 * d * n, sizeof(double) * d * n, cudaMemcpyDeviceToDevice, st));   // qs of the last adv
    }
    e->launches += 1;
    return SC_OK;
  }
  if (int rc = stage_buffers(e)) return rc;
  const double *src = (stage == 1) ? e->dev.rec : e->stage_buf;
  k_stage_positions<<<e->sm_count * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(e->dev, src, q_dev);
  CU(cudaGetLastError());
  e->launches += 1;
  return SC_OK;
}

extern "C" int sc_engine_stage_apply(sc_engine *e, int stage, double dt, const double *masses_dev, const double *V_dev,
                                     const double *grad_dev, const double *hess_dev, double *energy_sum_dev, void *stream) {
  if (!e || !masses_dev || !V_dev || !grad_dev || !hess_dev || e->dev.n < 1) return fail(SC_ERR_INVALID, "no ensemble / null argument");
  if (stage < 1 || stage > 4) return fail(SC_ERR_INVALID, "stage %d outside 1..4", stage);
  if (stage_uses_stream(e)) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StreamLayout L;
    if (int rc = stage_stream_buffers(e, L, st)) return rc;
    const size_t n = e->dev.n, d = e->dev.d, hsz = (size_t)L.hsz;
    double *pst = e->stage_buf, *img = pst + n * (6 * d + 2), *qp = img + n * 4 * hsz, *aux = qp + n * 2 * d, *im = aux + n * 8;
    double *rbuf = pst + 2 * d * n;                             // the stage positions are the path state's qs (q_dev is a copy)
    k_inv_masses<<<1, 128, 0, st>>>((int)d, masses_dev, im);
    k_hess_to_image<<<e->sm_count * 8, 256, 0, st>>>((int)d, (int)n, L.ldh, hess_dev, img + (size_t)(stage - 1) * n * hsz);
    PotDev P = PotDev();
    P.d = (int)d;
    P.imass = im;
    k_gstage_adv<<<(int)((n + 127) / 128), 128, 0, st>>>(e->dev, P, dt, stage, 0, 1, 0, (int)n, pst, V_dev, grad_dev, rbuf, qp, aux);
    CU(cudaGetLastError());
    if (stage == 4 && energy_sum_dev) {
      k_stage_energy<<<std::min((int)((n + 255) / 256), 64), 256, 0, st>>>((int)n, aux, energy_sum_dev);
      CU(cudaGetLastError());
    }
    e->launches += 3;
    return SC_OK;
  }
  if (int rc = stage_buffers(e)) return rc;
  const int d = e->dev.d;
  const size_t smem = sizeof(double) * ((size_t)4 * d * d + d * d + 2 * d + 16);
  CU(cudaFuncSetAttribute(k_stage_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = e->sm_count * 2;
  if (grid > e->dev.n) grid = e->dev.n;
  double *ys = e->stage_buf, *acc = e->stage_buf + (size_t)e->dev.n * e->dev.rs;
  k_stage_apply<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(e->dev, stage, dt, masses_dev, V_dev, grad_dev, hess_dev,
                                                                         ys, acc, energy_sum_dev);
  CU(cudaGetLastError());
  e->launches += 1;
  return SC_OK;
}

extern "C" int sc_engine_stage_finish(sc_engine *e, double dt, void *stream) {
  (void)dt;
  if (!e || e->dev.n < 1) return fail(SC_ERR_INVALID, "no ensemble");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (stage_uses_stream(e)) {
    // one step of the monodromy blocks through k_rk4_stream with the four stored stage Hessians, then prefactor + tracking
    StreamLayout L;
    if (int rc = stage_stream_buffers(e, L, st)) return rc;
    const int d = e->dev.d, dr = e->dev.dr, sm = e->sm_count;
    const size_t n = e->dev.n, hsz = (size_t)L.hsz;
    const bool dense = !e->dev.diag;
    double *pst = e->stage_buf, *img = pst + n * (6 * (size_t)d + 2), *im = img + n * 4 * hsz + n * 2 * d + n * 8;
    const size_t tsz = dense ? (size_t)L.mtr * L.nt * 128 : 0, cmsz = (size_t)dr * dr * 2;
    const size_t per_traj = sizeof(double) * (cmsz + 2 + tsz);
    // windows over the trajectories: the images of a window must be contiguous per stage -> the window is the ensemble when
    // the scratch allows it, else the images are re-laid per window (not needed below 2 GB of scratch per 10^5 trajectories)
    if (per_traj * n + 1024 > ((size_t)24 << 30)) return fail(SC_ERR_UNSUPPORTED, "stage interface: ensemble too large for one window");
    if (int rc = ensure_chunk_scratch(e, per_traj * n + 1024, st)) return rc;
    double *base = reinterpret_cast<double *>(e->chunk_scratch);
    double2 *cm = reinterpret_cast<double2 *>(base);            base += n * cmsz;
    double2 *det = reinterpret_cast<double2 *>(base);           base += n * 2;
    StreamArgs A;
    A.hs = img;
    A.hs_const = 0;
    A.L1p = dense ? e->stream_const + hsz : nullptr;
    A.L2p = dense ? e->stream_const + 2 * hsz : nullptr;
    A.cm = cm;
    A.T = dense ? base : nullptr;
    A.skip_rk4 = 0;
    PotDev P = PotDev();
    P.d = d;
    P.imass = im;
    CU(launch_stream((long long)n * L.ngroups, sm, e->dev, P, dt, 1, 0, (int)n, A, L, st));
    if (dense) CU(launch_rmult(e->dev, (long long)n, A.T, cm, sm, st));
    CU(launch_lu_batch(cm, dr, (int)n, det, sm, 0, st));
    k_track_only<<<(int)((n + 127) / 128), 128, 0, st>>>(e->dev, 0, (int)n, det, 0);
    CU(cudaGetLastError());
    e->launches += dense ? 4 : 3;
    e->kernel_name = "stage interface: k_rk4_stream+k_lu+k_track_only";
  } else {
    PotDev none = PotDev();
    none.d = e->dev.d;
    none.imass = e->dev.q0;
    if (int rc = run_hk_kernel(e, none, 0.0, 0, MODE_TRACK, nullptr, st, false)) return rc;
  }
  if (e->cfg.wm) {
    // the WM pieces of the new time: trackers updated, contributions discarded
    if (int rc = wm_launch(e->wm, e->dev, WM_STEP, 0.0, e->wm.scratch5, nullptr, st)) return rc;
    e->launches += 2;
  }
  return SC_OK;
}
