// TEST INFRASTRUCTURE: runs the __host__ __device__ Walton-Manolopoulos trajectory routine of the product
// (semiclassical_b200/csrc/sc_wm.cuh) on the HOST with one "thread" per group, so that its formulas can be
// checked against the oracle / reference goldens without a GPU.  Not part of the product path.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../semiclassical_b200/csrc/sc_wm.cuh"

using namespace sc;

// y: reference layout (2d+4d^2+1, n); zi (2d, n); c: n complex; signs: (3, n) rows C, detA, detM
// out4: sums of the contributions (no 1/N); det_out: (n, 4) = detA.re, detA.im, detM.re, detM.im
extern "C" int wm_emul(int d, int dr, int n, const double *G0, const double *Gi, const double *Gt, const double *iGi0,
                       const double *iG0, const double *U, const double *q0, const double *p0, const double *n1,
                       double alpha, double beta, double pref, const double *y, const double *zi, const double *probi,
                       const double *c, const double *signs, double *out4, double *det_out) {
  const int d2 = d * d;
  std::vector<double> GiG(d2), Cqq(d2);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += G0[i * d + k] * iGi0[k * d + j];
      GiG[i * d + j] = s;
    }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += GiG[i * d + k] * G0[k * d + j];
      Cqq[i * d + j] = G0[i * d + j] - s;
    }
  EngDev E = EngDev();
  E.d = d; E.dr = dr; E.n = n;
  E.qps = (2 * d + 1 + 1) & ~1;
  E.rs = E.qps + 4 * d2;
  std::vector<double> rec((size_t)n * E.rs, 0.0), zt((size_t)n * 2 * d), winv(n), sA(n), sM(n), s0(n);
  std::vector<double2> cc(n), prevA(n), prevM(n);
  const double inv2pid = std::pow(2.0 * M_PI, -(double)d);
  for (int t = 0; t < n; ++t) {
    double *r = rec.data() + (size_t)t * E.rs;
    for (int k = 0; k < 2 * d; ++k) { r[k] = y[(size_t)k * n + t]; zt[(size_t)t * 2 * d + k] = zi[(size_t)k * n + t]; }
    r[2 * d] = y[(size_t)(2 * d + 4 * d2) * n + t];
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) {
        r[E.qps + a * 2 * d + b] = y[(size_t)(2 * d + a * d + b) * n + t];                    // Mqq
        r[E.qps + a * 2 * d + d + b] = y[(size_t)(2 * d + d2 + a * d + b) * n + t];           // Mqp
        r[E.qps + 2 * d2 + a * 2 * d + b] = y[(size_t)(2 * d + 2 * d2 + a * d + b) * n + t];  // Mpq
        r[E.qps + 2 * d2 + a * 2 * d + d + b] = y[(size_t)(2 * d + 3 * d2 + a * d + b) * n + t];
      }
    winv[t] = inv2pid / probi[t];
    cc[t] = make_double2(c[2 * t], c[2 * t + 1]);
    s0[t] = signs[t];
  }
  E.rec = rec.data();
  E.zt = zt.data();
  E.c = cc.data();
  E.sign = s0.data();
  WMDev W = WMDev();
  W.d = d; W.dr = dr;
  W.G0 = G0; W.Gi = Gi; W.Gt = Gt; W.iGi0 = iGi0; W.iG0 = iG0; W.GiG = GiG.data(); W.Cqq = Cqq.data(); W.U = U;
  W.q0 = q0; W.p0 = p0; W.n1 = n1;
  W.alpha = alpha; W.beta = beta; W.pref = pref;
  W.prevA = prevA.data(); W.prevM = prevM.data(); W.signA = sA.data(); W.signM = sM.data(); W.winv = winv.data();
  const WMLayout L = make_wm_layout(d, dr);
  std::vector<double2> ws(L.total);
  double acc[4] = {0, 0, 0, 0};
  for (int t = 0; t < n; ++t) wm_trajectory<1>(E, W, L, ws.data(), t, WM_INIT, 0, 0, acc);
  for (int t = 0; t < n; ++t) {
    det_out[4 * t + 0] = prevA[t].x; det_out[4 * t + 1] = prevA[t].y;
    det_out[4 * t + 2] = prevM[t].x; det_out[4 * t + 3] = prevM[t].y;
    sA[t] = signs[n + t];
    sM[t] = signs[2 * n + t];
  }
  for (int t = 0; t < n; ++t) wm_trajectory<1>(E, W, L, ws.data(), t, WM_CORR, 0, 0, acc);
  for (int k = 0; k < 4; ++k) out4[k] = acc[k];
  return 0;
}
