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

static int stage_buffers(sc_engine *e) {
  if (e->stage_buf) return SC_OK;
  CU(e->ens.alloc((size_t)2 * e->dev.n * e->dev.rs, &e->stage_buf));
  return SC_OK;
}

extern "C" int sc_engine_stage_positions(sc_engine *e, int stage, double dt, double *q_dev, void *stream) {
  (void)dt;
  if (!e || !q_dev || e->dev.n < 1) return fail(SC_ERR_INVALID, "no ensemble / null argument");
  if (stage < 1 || stage > 4) return fail(SC_ERR_INVALID, "stage %d outside 1..4", stage);
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
  PotDev none = PotDev();
  none.d = e->dev.d;
  none.imass = e->dev.q0;
  if (int rc = run_hk_kernel(e, none, 0.0, 0, MODE_TRACK, nullptr, st, false)) return rc;
  if (e->cfg.wm) {
    // the WM pieces of the new time: trackers updated, contributions discarded
    if (int rc = wm_launch(e->wm, e->dev, WM_STEP, 0.0, e->wm.scratch5, nullptr, st)) return rc;
    e->launches += 2;
  }
  return SC_OK;
}
