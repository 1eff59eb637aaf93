// sc_potentials.cuh -- batched energy / gradient / Hessian kernels behind potential.harmonic_approximation(r)
// (potentials.py:136-162, 329-355, 553-593, 669-699).  Batch-last layout: r (d, n), grad (d, n), hess (d, d, n),
// one thread per geometry so that every global access is coalesced over the trajectory index.
#pragma once
#include "sc_device.cuh"

namespace sc {

// separable per-mode potentials (Morse/AS, 1-D Herman-Kluk model)
__global__ void k_pot_separable(PotDev P, int n, const double *__restrict__ r, double *__restrict__ V,
                                double *__restrict__ grad, double *__restrict__ hess) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = P.d;
  double v = 0.0;
  for (int k = 0; k < d; ++k) {
    const double x = r[(size_t)k * n + i];
    double g, h;
    if (P.type == POT_MORSE) {
      if (P.all_harmonic) {
        const double w2 = P.omega[k] * P.omega[k];
        v += 0.5 * w2 * x * x; g = w2 * x; h = w2;
      } else {
        const double a = P.a[k], D = P.D[k], e = exp(-a * x);
        v += D * (1.0 - e) * (1.0 - e);
        g = 2.0 * a * D * e * (1.0 - e);
        h = 2.0 * a * a * D * e * (2.0 * e - 1.0);
      }
    } else {
      const double eps = P.eps[k], b = P.b[k], e1 = exp(-b * x), e2 = exp(-2.0 * b * x);
      v += eps / (2.0 * b * b) * (1.0 - e1) * (1.0 - e1) + (1.0 - eps) * 0.5 * x * x;
      g = eps / b * (e1 - e2) + (1.0 - eps) * x;
      h = eps * (2.0 * e2 - e1) + (1.0 - eps);
    }
    if (grad) grad[(size_t)k * n + i] = g;
    if (hess)
      for (int l = 0; l < d; ++l) hess[((size_t)k * d + l) * n + i] = (l == k) ? h : 0.0;
  }
  V[i] = v - P.origin;
}

// molecular harmonic expansion (potentials.py:581-593): thread per (geometry), rows of H0 broadcast from L1
__global__ void k_pot_harmonic(PotDev P, int n, const double *__restrict__ r, double *__restrict__ V,
                               double *__restrict__ grad, double *__restrict__ hess) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = P.d;
  double v = P.e0 - P.origin;
  for (int a = 0; a < d; ++a) {
    double hd = 0.0;
    for (int b = 0; b < d; ++b) hd += P.hess0[a * d + b] * (r[(size_t)b * n + i] - P.pos0[b]);
    const double dra = r[(size_t)a * n + i] - P.pos0[a];
    v += dra * P.grad0[a] + 0.5 * dra * hd;
    if (grad) grad[(size_t)a * n + i] = P.grad0[a] + hd;
    if (hess)
      for (int b = 0; b < d; ++b) hess[((size_t)a * d + b) * n + i] = P.hess0[a * d + b];
  }
  V[i] = v;
}

// rotated Morse: x = Q r  (dense-path fixture)
__global__ void k_pot_rotated(PotDev P, int n, const double *__restrict__ x, double *__restrict__ V,
                              double *__restrict__ grad, double *__restrict__ hess, double *__restrict__ work) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = P.d;
  double *gi = work + (size_t)i, *hi = work + (size_t)d * n + i;  // (d, n) each, batch-last
  double v = 0.0;
  for (int k = 0; k < d; ++k) {
    double s = 0.0;
    for (int a = 0; a < d; ++a) s += P.Q[a * d + k] * x[(size_t)a * n + i];
    double g, h;
    if (P.all_harmonic) {
      const double w2 = P.omega[k] * P.omega[k];
      v += 0.5 * w2 * s * s; g = w2 * s; h = w2;
    } else {
      const double a_ = P.a[k], D = P.D[k], e = exp(-a_ * s);
      v += D * (1.0 - e) * (1.0 - e);
      g = 2.0 * a_ * D * e * (1.0 - e);
      h = 2.0 * a_ * a_ * D * e * (2.0 * e - 1.0);
    }
    gi[(size_t)k * n] = g;
    hi[(size_t)k * n] = h;
  }
  V[i] = v - P.origin;
  for (int a = 0; a < d; ++a) {
    if (grad) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += P.Q[a * d + k] * gi[(size_t)k * n];
      grad[(size_t)a * n + i] = s;
    }
    if (hess)
      for (int b = 0; b < d; ++b) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += P.Q[a * d + k] * hi[(size_t)k * n] * P.Q[b * d + k];
        hess[((size_t)a * d + b) * n + i] = s;
      }
  }
}

}  // namespace sc

#include "sc_gdml.cuh"

namespace sc {

// returns 0 on launch, 1 if the potential type has no batched kernel
static int launch_potential_eval(const PotDev &P, int n, const double *r, double *V, double *grad, double *hess,
                                 cudaStream_t st) {
  const int threads = 128, blocks = (n + threads - 1) / threads;
  switch (P.type) {
    case POT_MORSE:
    case POT_NONHARMONIC:
      k_pot_separable<<<blocks, threads, 0, st>>>(P, n, r, V, grad, hess);
      return 0;
    case POT_HARMONIC:
      k_pot_harmonic<<<blocks, threads, 0, st>>>(P, n, r, V, grad, hess);
      return 0;
    case POT_ROTATED_MORSE: {
      double *work = nullptr;
      if (cudaMallocAsync(&work, sizeof(double) * 2 * (size_t)P.d * n, st) != cudaSuccess) return 1;
      k_pot_rotated<<<blocks, threads, 0, st>>>(P, n, r, V, grad, hess, work);
      cudaFreeAsync(work, st);
      return 0;
    }
    case POT_GDML:
      return launch_gdml_eval(P, n, r, V, grad, hess, st);
    default:
      return 1;
  }
}

}  // namespace sc
