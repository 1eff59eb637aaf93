// sc_mma.cuh -- FP64 tensor-core primitive and the separable-potential device functions shared by the column pipelines.
//
// tcgen05/TMEM has no f64 kind, so the tensor path of this problem is warp-level mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4):
// A (8 x 4, row) one element per lane at (row = lane / 4, k = lane % 4), B (4 x 8, col) one element per lane at
// (k = lane % 4, n = lane / 4), C (8 x 8) two elements per lane at (row = lane / 4, cols 2 (lane % 4), + 1).
// (The round-1 general kernel k_hk_mma -- one 210-KB CTA per trajectory with the potential, the prefactor assembly and the LU
// inside -- lived here; the dense column pipeline of sc_stream.cuh replaced it in round 2.)
#pragma once
#include "sc_kernels.cuh"

namespace sc {

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// separable potentials without side effects: value, gradient and Hessian diagonal of mode t at r
__device__ __forceinline__ double pot_local_vals(const PotDev &P, int t, double r, double &gout, double &hd) {
  double v;
  if (P.type == POT_MORSE) {
    if (P.all_harmonic) {
      const double w2 = P.omega[t] * P.omega[t];
      v = 0.5 * w2 * r * r;
      gout = w2 * r;
      hd = w2;
    } else {
      const double a = P.a[t], D = P.D[t];
      const double e = exp(-a * r);
      v = D * (1.0 - e) * (1.0 - e);
      gout = 2.0 * a * D * e * (1.0 - e);
      hd = 2.0 * a * a * D * e * (2.0 * e - 1.0);
    }
  } else {
    const double eps = P.eps[t], b = P.b[t];
    const double e1 = exp(-b * r), e2 = exp(-2.0 * b * r);
    v = eps / (2.0 * b * b) * (1.0 - e1) * (1.0 - e1) + (1.0 - eps) * 0.5 * r * r;
    gout = eps / b * (e1 - e2) + (1.0 - eps) * r;
    hd = eps * (2.0 * e2 - e1) + (1.0 - eps);
  }
  if (t == 0) v -= P.origin;
  return v;
}

}  // namespace sc
