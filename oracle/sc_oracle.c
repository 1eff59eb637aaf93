/*
 * sc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (C99 + OpenMP over trajectories) CPU restatement of the reference's Herman-Kluk /
 * Walton-Manolopoulos propagation path, used only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the CPU baseline.  The product path
 * (semiclassical_b200/) never links or calls it.
 *
 * Pinned against golden vectors produced by the unmodified reference (oracle/make_golden.py ->
 * tests/golden/, checked in tests/test_oracle.py): parity is PINNED, not "unpinned".
 *
 * Reference lines restated (paths relative to /root/reference/semiclassical):
 *   rk4_step            propagators.py:86-119
 *   eom_rhs             propagators.py:313-383
 *   hk_prefactor        propagators.py:951-1004
 *   track_sign          propagators.py:1006-1052
 *   cs_overlap          propagators.py:181-240  (ket = single state q0,p0)
 *   hk contributions    propagators.py:784-911
 *   wm_prefactor        propagators.py:1132-1389
 *   wm contributions    propagators.py:1577-1719
 *   potentials          potentials.py:63-134 (1-D HK model), 265-327 (Morse/AS), 581-593 (molecular harmonic)
 *   gdml_eval           gdml_predictor.py:140-250
 *
 * Layout conventions of the C interface: matrices row-major; ensemble arrays batch-last like the
 * reference (zi is (2d, n): zi[k*n + traj]); y_out is the reference's (2d+4d^2+1, n) state.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cplx;

enum { POT_MORSE = 0, POT_HARMONIC = 2, POT_NONHARMONIC = 3, POT_ROTATED_MORSE = 4, POT_GDML = 5 };

typedef struct {
  int type, d;
  const double *masses, *nac;
  /* Morse / AS */
  const double *omega, *a, *D;
  int all_harmonic;
  /* molecular harmonic expansion */
  const double *pos0, *grad0, *hess0;
  double energy0, origin;
  /* 1-D Herman-Kluk test potential (per mode) */
  const double *eps, *b;
  /* rotation x = Q r (d x d, row-major) around a Morse potential */
  const double *Q;
  /* sGDML */
  int n_atoms, n_train, n_desc;
  const double *xs_train, *jx_alphas; /* (M, D) row-major */
  double sig, c, std;
} sc_oracle_potential;

typedef struct {
  int d, dr;
  const cplx *sqGi, *isqGi, *sqGt, *isqGt; /* d x d */
  const cplx *U;                            /* d x dr */
  const double *oi0_A, *oi0_B, *oi0_C;      /* overlap <.,Gi|.,G0>: Gi iGij Gj, iGij, Gj iGij */
  double oi0_fac;
  const double *ot0_A, *ot0_B, *ot0_C;      /* overlap <.,Gt|.,G0> */
  double ot0_fac;
  const double *q0, *p0;
  const double *Gamma_0, *Gamma_i, *Gamma_t, *iGi0;
  /* Walton-Manolopoulos only */
  double alpha, beta;
  const double *iGamma_0;
  double detG0, detGi, detGt, detGi0;
} sc_oracle_consts;

/* ------------------------------------------------------------------ potentials ------------- */

static void morse_eval(const sc_oracle_potential *P, const double *r, double *V, double *g, double *h) {
  const int d = P->d;
  double v = 0.0;
  memset(h, 0, sizeof(double) * d * d);
  for (int k = 0; k < d; ++k) {
    if (P->all_harmonic) {
      const double w2 = P->omega[k] * P->omega[k];
      v += 0.5 * w2 * r[k] * r[k];
      g[k] = w2 * r[k];
      h[k * d + k] = w2;
    } else {
      const double a = P->a[k], D = P->D[k];
      const double e = exp(-a * r[k]);
      v += D * (1.0 - e) * (1.0 - e);
      g[k] = 2 * a * D * e * (1.0 - e);
      h[k * d + k] = 2 * a * a * D * e * (2 * e - 1.0);
    }
  }
  *V = v;
}

static void gdml_eval(const sc_oracle_potential *P, const double *r, double *E, double *grad, double *hess);

static void pot_eval(const sc_oracle_potential *P, const double *r, double *V, double *g, double *h) {
  const int d = P->d;
  switch (P->type) {
    case POT_MORSE:
      morse_eval(P, r, V, g, h);
      break;
    case POT_NONHARMONIC: {
      double v = 0.0;
      memset(h, 0, sizeof(double) * d * d);
      for (int k = 0; k < d; ++k) {
        const double eps = P->eps[k], b = P->b[k];
        const double e1 = exp(-b * r[k]), e2 = exp(-2 * b * r[k]);
        v += eps / (2 * b * b) * (1.0 - e1) * (1.0 - e1) + (1 - eps) * 0.5 * r[k] * r[k];
        g[k] = eps / b * (e1 - e2) + (1 - eps) * r[k];
        h[k * d + k] = eps * (2 * e2 - e1) + (1 - eps);
      }
      *V = v;
      break;
    }
    case POT_HARMONIC: {
      double *dr = (double *)malloc(sizeof(double) * d);
      double v = P->energy0;
      for (int i = 0; i < d; ++i) dr[i] = r[i] - P->pos0[i];
      for (int i = 0; i < d; ++i) {
        double hd = 0.0;
        for (int j = 0; j < d; ++j) hd += P->hess0[i * d + j] * dr[j];
        g[i] = P->grad0[i] + hd;
        v += dr[i] * P->grad0[i] + 0.5 * dr[i] * hd;
      }
      memcpy(h, P->hess0, sizeof(double) * d * d);
      *V = v - P->origin;
      free(dr);
      break;
    }
    case POT_ROTATED_MORSE: {
      /* V'(x) = V(Q^T x), grad' = Q grad, hess' = Q hess Q^T */
      double *rr = (double *)malloc(sizeof(double) * (2 * d + d * d));
      double *gi = rr + d, *hi = gi + d;
      const double *Q = P->Q;
      for (int k = 0; k < d; ++k) {
        double s = 0.0;
        for (int i = 0; i < d; ++i) s += Q[i * d + k] * r[i];
        rr[k] = s;
      }
      morse_eval(P, rr, V, gi, hi);
      for (int i = 0; i < d; ++i) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += Q[i * d + k] * gi[k];
        g[i] = s;
      }
      for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
          double s = 0.0;
          for (int k = 0; k < d; ++k) s += Q[i * d + k] * hi[k * d + k] * Q[j * d + k];
          h[i * d + j] = s;
        }
      free(rr);
      break;
    }
    case POT_GDML:
      gdml_eval(P, r, V, g, h);
      *V -= P->origin;
      break;
    default:
      *V = NAN;
  }
}

/* sGDML energy / gradient / Hessian for one geometry (gdml_predictor.py:140-250) */
static void gdml_eval(const sc_oracle_potential *P, const double *r, double *E, double *grad, double *hess) {
  const int N = P->n_atoms, M = P->n_train, D = P->n_desc, X = 3 * N;
  const double q = sqrt(5.0) / P->sig;
  double *xs = (double *)calloc((size_t)D * (2 + X) + (size_t)M * (4 + 2 * X) + (size_t)D, sizeof(double));
  double *gx = xs + D;              /* dE/dx_desc (D) */
  double *J = gx + D;               /* Jacobian (D, X) */
  double *xd = J + (size_t)D * X;   /* scratch row x - x_m (D) */
  double *xn = xd + D;              /* |x - x_m| (M) */
  double *XA = xn + M;              /* (M) */
  double *ef = XA + M;              /* exp factor (M) */
  double *k1 = ef + M;              /* ef (1+q r)/q^2 (M) */
  double *XJ = k1 + M;              /* (M, X) */
  double *AJ = XJ + (size_t)M * X;  /* (M, X) */
  int *pi_ = (int *)malloc(sizeof(int) * 2 * D), *pj_ = pi_ + D;
  {
    int n = 0;
    for (int i = 1; i < N; ++i)
      for (int j = 0; j < i; ++j) { pi_[n] = i; pj_[n] = j; ++n; }
  }
  for (int n = 0; n < D; ++n) {
    const int i = pi_[n], j = pj_[n];
    double dx[3], r2 = 0;
    for (int u = 0; u < 3; ++u) { dx[u] = r[3 * i + u] - r[3 * j + u]; r2 += dx[u] * dx[u]; }
    xs[n] = 1.0 / sqrt(r2);
    const double x3 = xs[n] * xs[n] * xs[n];
    for (int u = 0; u < 3; ++u) {
      J[(size_t)n * X + 3 * i + u] = -x3 * dx[u];
      J[(size_t)n * X + 3 * j + u] = x3 * dx[u];
    }
  }
  double en = 0.0;
  memset(gx, 0, sizeof(double) * D);
  for (int m = 0; m < M; ++m) {
    const double *xt = P->xs_train + (size_t)m * D, *A = P->jx_alphas + (size_t)m * D;
    double n2 = 0, xa = 0;
    for (int n = 0; n < D; ++n) { xd[n] = xs[n] - xt[n]; n2 += xd[n] * xd[n]; xa += xd[n] * A[n]; }
    xn[m] = sqrt(n2);
    XA[m] = xa;
    ef[m] = 1.0 / 3.0 * q * q * q * q * exp(-q * xn[m]);
    k1[m] = ef[m] * (1.0 + q * xn[m]) / (q * q);
    en += k1[m] * xa;
    for (int n = 0; n < D; ++n) gx[n] += k1[m] * A[n] - ef[m] * xa * xd[n];
    for (int x = 0; x < X; ++x) {
      double s1 = 0, s2 = 0;
      for (int n = 0; n < D; ++n) { s1 += xd[n] * J[(size_t)n * X + x]; s2 += A[n] * J[(size_t)n * X + x]; }
      XJ[(size_t)m * X + x] = s1;
      AJ[(size_t)m * X + x] = s2;
    }
  }
  *E = en * P->std + P->c;
  for (int x = 0; x < X; ++x) {
    double s = 0;
    for (int n = 0; n < D; ++n) s += gx[n] * J[(size_t)n * X + x];
    grad[x] = s * P->std;
  }
  double sumefxa = 0;
  for (int m = 0; m < M; ++m) sumefxa += ef[m] * XA[m];
  for (int x = 0; x < X; ++x)
    for (int y = 0; y < X; ++y) {
      double s = 0, jj = 0;
      for (int m = 0; m < M; ++m) {
        const double xjx = XJ[(size_t)m * X + x], xjy = XJ[(size_t)m * X + y];
        s += ef[m] * XA[m] * q / xn[m] * xjx * xjy - ef[m] * (AJ[(size_t)m * X + x] * xjy + xjx * AJ[(size_t)m * X + y]);
      }
      for (int n = 0; n < D; ++n) jj += J[(size_t)n * X + x] * J[(size_t)n * X + y];
      hess[x * X + y] = s - sumefxa * jj;
    }
  /* second derivative of the descriptor: h1, h2 scatter terms */
  for (int n = 0; n < D; ++n) {
    const int k = pi_[n], l = pj_[n];
    double dx[3];
    for (int u = 0; u < 3; ++u) dx[u] = r[3 * k + u] - r[3 * l + u];
    const double x5 = pow(xs[n], 5), x3 = xs[n] * xs[n] * xs[n];
    const double h2 = -gx[n] * x3;
    for (int u = 0; u < 3; ++u) {
      for (int v = 0; v < 3; ++v) {
        const double h1 = 3 * gx[n] * x5 * dx[u] * dx[v];
        hess[(3 * k + u) * X + 3 * l + v] -= h1;
        hess[(3 * l + u) * X + 3 * k + v] -= h1;
        hess[(3 * k + u) * X + 3 * k + v] += h1;
        hess[(3 * l + u) * X + 3 * l + v] += h1;
      }
      hess[(3 * k + u) * X + 3 * l + u] -= h2;
      hess[(3 * l + u) * X + 3 * k + u] -= h2;
      hess[(3 * k + u) * X + 3 * k + u] += h2;
      hess[(3 * l + u) * X + 3 * l + u] += h2;
    }
  }
  for (int x = 0; x < X * X; ++x) hess[x] *= P->std;
  free(pi_);
  free(xs);
}

/* batched potential interface: r (d, n) batch-last -> V (n), grad (d, n), hess (d, d, n) */
int sc_oracle_potential_eval(const sc_oracle_potential *P, int n, const double *r, double *V, double *grad,
                             double *hess) {
  const int d = P->d;
#pragma omp parallel
  {
    double *buf = (double *)malloc(sizeof(double) * (2 * d + d * d));
    double *rr = buf, *g = buf + d, *h = g + d;
#pragma omp for schedule(static)
    for (int t = 0; t < n; ++t) {
      for (int k = 0; k < d; ++k) rr[k] = r[(size_t)k * n + t];
      double v;
      pot_eval(P, rr, &v, g, h);
      V[t] = v;
      for (int k = 0; k < d; ++k) grad[(size_t)k * n + t] = g[k];
      for (int k = 0; k < d * d; ++k) hess[(size_t)k * n + t] = h[k];
    }
    free(buf);
  }
  return 0;
}

/* ------------------------------------------------------------------ small dense helpers ---- */

/* determinant of an n x n complex matrix by LU with partial pivoting (destroys a) */
static cplx cdet(cplx *a, int n) {
  cplx det = 1.0;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = cabs(a[k * n + k]);
    for (int i = k + 1; i < n; ++i) {
      const double v = cabs(a[i * n + k]);
      if (v > best) { best = v; p = i; }
    }
    if (best == 0.0) return 0.0;
    if (p != k) {
      for (int j = 0; j < n; ++j) { cplx t = a[k * n + j]; a[k * n + j] = a[p * n + j]; a[p * n + j] = t; }
      det = -det;
    }
    const cplx piv = a[k * n + k];
    det *= piv;
    for (int i = k + 1; i < n; ++i) {
      const cplx f = a[i * n + k] / piv;
      for (int j = k + 1; j < n; ++j) a[i * n + j] -= f * a[k * n + j];
    }
  }
  return det;
}

/* inverse by Gauss-Jordan with partial pivoting: a (n x n) -> inv (n x n); destroys a */
static void cinv(cplx *a, cplx *inv, int n) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) inv[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = cabs(a[k * n + k]);
    for (int i = k + 1; i < n; ++i) {
      const double v = cabs(a[i * n + k]);
      if (v > best) { best = v; p = i; }
    }
    if (p != k)
      for (int j = 0; j < n; ++j) {
        cplx t = a[k * n + j]; a[k * n + j] = a[p * n + j]; a[p * n + j] = t;
        t = inv[k * n + j]; inv[k * n + j] = inv[p * n + j]; inv[p * n + j] = t;
      }
    const cplx ipiv = 1.0 / a[k * n + k];
    for (int j = 0; j < n; ++j) { a[k * n + j] *= ipiv; inv[k * n + j] *= ipiv; }
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      const cplx f = a[i * n + k];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) { a[i * n + j] -= f * a[k * n + j]; inv[i * n + j] -= f * inv[k * n + j]; }
    }
  }
}

/* C (m x n) = A (m x k) * B (k x n), complex, row-major; tb: use B^T (B given as n x k) */
static void cgemm(int m, int n, int k, const cplx *A, const cplx *B, int tb, cplx *C) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      cplx s = 0.0;
      if (tb)
        for (int l = 0; l < k; ++l) s += A[i * k + l] * B[j * k + l];
      else
        for (int l = 0; l < k; ++l) s += A[i * k + l] * B[l * n + j];
      C[i * n + j] = s;
    }
}

/* x^T A y for a real d x d matrix and real vectors */
static double quad(const double *A, const double *x, const double *y, int d) {
  double s = 0.0;
  for (int i = 0; i < d; ++i) {
    double t = 0.0;
    for (int j = 0; j < d; ++j) t += A[i * d + j] * y[j];
    s += x[i] * t;
  }
  return s;
}

/* x^T A y for a complex d x d matrix, x and y complex (no conjugation) */
static cplx cquad(const cplx *A, const cplx *x, const cplx *y, int d) {
  cplx s = 0.0;
  for (int i = 0; i < d; ++i) {
    cplx t = 0.0;
    for (int j = 0; j < d; ++j) t += A[i * d + j] * y[j];
    s += x[i] * t;
  }
  return s;
}

/* ------------------------------------------------------------------ per-trajectory state --- */

typedef struct {
  double *y;   /* q(d) p(d) Mqq(d^2) Mqp Mpq Mpp S  -- the reference's row order */
  cplx c2;     /* det of the prefactor matrix */
  cplx c;      /* principal sqrt(c2) */
  /* sign trackers: previous z and sign for "prefactorC", "detA", "detM" */
  cplx prev[3];
  double sign[3];
  int tracked[3];
  /* WM per-trajectory quantities */
  cplx *Rqq, *RQQ, *RqQ, *Pq, *PQ;
  cplx gamma, detA, detM;
} traj_t;

static int ylen(int d) { return 2 * d + 4 * d * d + 1; }

/* dy/dt (propagators.py:313-383); returns T+V of this stage point */
static double eom_rhs(const sc_oracle_potential *P, const double *y, double *dy, double *g, double *h) {
  const int d = P->d, d2 = d * d;
  const double *q = y, *p = y + d, *Mqq = y + 2 * d, *Mqp = Mqq + d2, *Mpq = Mqp + d2, *Mpp = Mpq + d2;
  double *Dq = dy, *Dp = dy + d, *DMqq = dy + 2 * d, *DMqp = DMqq + d2, *DMpq = DMqp + d2, *DMpp = DMpq + d2;
  double V, tkin = 0.0;
  pot_eval(P, q, &V, g, h);
  for (int a = 0; a < d; ++a) {
    const double im = 1.0 / P->masses[a];
    Dq[a] = p[a] * im;
    Dp[a] = -g[a];
    tkin += 0.5 * p[a] * p[a] * im;
    for (int b = 0; b < d; ++b) {
      DMqq[a * d + b] = Mpq[a * d + b] * im;
      DMqp[a * d + b] = Mpp[a * d + b] * im;
      double s1 = 0.0, s2 = 0.0;
      for (int k = 0; k < d; ++k) {
        s1 += h[a * d + k] * Mqq[k * d + b];
        s2 += h[a * d + k] * Mqp[k * d + b];
      }
      DMpq[a * d + b] = -s1;
      DMpp[a * d + b] = -s2;
    }
  }
  dy[2 * d + 4 * d2] = tkin - V;
  return tkin + V;
}

/* classical RK4 (propagators.py:114-119); *etot = (T+V) at the 4th stage point */
static void rk4_step(const sc_oracle_potential *P, double *y, double hstep, double *work, double *etot) {
  const int n = ylen(P->d), d = P->d;
  double *k1 = work, *k2 = k1 + n, *k3 = k2 + n, *k4 = k3 + n, *ys = k4 + n, *g = ys + n, *h = g + d;
  eom_rhs(P, y, k1, g, h);
  for (int i = 0; i < n; ++i) ys[i] = y[i] + 0.5 * hstep * k1[i];
  eom_rhs(P, ys, k2, g, h);
  for (int i = 0; i < n; ++i) ys[i] = y[i] + 0.5 * hstep * k2[i];
  eom_rhs(P, ys, k3, g, h);
  for (int i = 0; i < n; ++i) ys[i] = y[i] + hstep * k3[i];
  *etot = eom_rhs(P, ys, k4, g, h);
  for (int i = 0; i < n; ++i) y[i] = y[i] + hstep / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

/* sqrt branch tracking (propagators.py:1035-1051) */
static void track_sign(traj_t *T, int key, cplx z) {
  if (!T->tracked[key]) {
    T->tracked[key] = 1;
    T->sign[key] = 1.0;
    T->prev[key] = z;
  }
  const cplx z1 = T->prev[key];
  if (creal(z1) < 0 && creal(z) < 0 && cimag(z1) * cimag(z) < 0) T->sign[key] = -T->sign[key];
  T->prev[key] = z;
}

/* HK prefactor: mat = 1/2 (sqGt Mqq isqGi + isqGt Mpp sqGi - i sqGt Mqp sqGi + i isqGt Mpq isqGi),
   projected with U, determinant, principal square root (propagators.py:959-1004) */
static void hk_prefactor(const sc_oracle_consts *K, traj_t *T, cplx *w) {
  const int d = K->d, dr = K->dr, d2 = d * d;
  const double *Mqq = T->y + 2 * d, *Mqp = Mqq + d2, *Mpq = Mqp + d2, *Mpp = Mpq + d2;
  cplx *M = w, *t1 = M + d2, *t2 = t1 + d2, *mat = t2 + d2, *sub = mat + d2;
  memset(mat, 0, sizeof(cplx) * d2);
  const double *blocks[4] = {Mqq, Mpp, Mqp, Mpq};
  const cplx *left[4] = {K->sqGt, K->isqGt, K->sqGt, K->isqGt};
  const cplx *right[4] = {K->isqGi, K->sqGi, K->sqGi, K->isqGi};
  const cplx coef[4] = {1.0, 1.0, -I, I};
  for (int b = 0; b < 4; ++b) {
    for (int i = 0; i < d2; ++i) M[i] = blocks[b][i];
    cgemm(d, d, d, left[b], M, 0, t1);
    cgemm(d, d, d, t1, right[b], 0, t2);
    for (int i = 0; i < d2; ++i) mat[i] += 0.5 * coef[b] * t2[i];
  }
  /* sub = U^T mat U */
  for (int a = 0; a < dr; ++a)
    for (int j = 0; j < d; ++j) {
      cplx s = 0.0;
      for (int i = 0; i < d; ++i) s += K->U[i * dr + a] * mat[i * d + j];
      t1[a * d + j] = s;
    }
  for (int a = 0; a < dr; ++a)
    for (int b = 0; b < dr; ++b) {
      cplx s = 0.0;
      for (int j = 0; j < d; ++j) s += t1[a * d + j] * K->U[j * dr + b];
      sub[a * dr + b] = s;
    }
  T->c2 = cdet(sub, dr);
  T->c = csqrt(T->c2);
  track_sign(T, 0, T->c2);
}

/* <q,p,Gbra | q0,p0,G0> with precomputed A = Gi iGij Gj, B = iGij, C = Gj iGij (propagators.py:230-237) */
static cplx cs_overlap(const double *A, const double *B, const double *C, double fac, const double *q,
                       const double *p, const double *q0, const double *p0, int d, double *tmp) {
  double *dq = tmp, *dp = tmp + d;
  double pjdq = 0.0;
  for (int i = 0; i < d; ++i) { dq[i] = q0[i] - q[i]; dp[i] = p0[i] - p[i]; pjdq += p0[i] * dq[i]; }
  const double re = -0.5 * quad(A, dq, dq, d) - 0.5 * quad(B, dp, dp, d);
  const double im = -pjdq + quad(C, dq, dp, d);
  return fac * cexp(re + I * im);
}

/* HK: per-trajectory contributions to C_auto and k_ic without the e^{itE0} phase and without 1/ntraj
   (propagators.py:784-807, 868-909) */
static void hk_contrib(const sc_oracle_potential *P, const sc_oracle_consts *K, const traj_t *T, const double *zi,
                       double probi, cplx *cauto, cplx *kic, double *tmp) {
  const int d = K->d;
  const double *q = zi, *p = zi + d, *Q = T->y, *Pm = T->y + d;
  const double S = T->y[2 * d + 4 * d * d];
  const cplx vi = cs_overlap(K->oi0_A, K->oi0_B, K->oi0_C, K->oi0_fac, q, p, K->q0, K->p0, d, tmp);
  const cplx vt = cs_overlap(K->ot0_A, K->ot0_B, K->ot0_C, K->ot0_fac, Q, Pm, K->q0, K->p0, d, tmp);
  const cplx cq = conj(vt) * vi * (T->sign[0] * T->c) * cexp(I * S);
  const double w = 1.0 / (probi * pow(2 * M_PI, d));
  *cauto = cq * w;
  /* n1 = -tau1/m (constant NAC), n2 = 0 */
  double *n1 = tmp, *dq = tmp + d, *dQ = tmp + 2 * d, *PI = tmp + 3 * d, *pi = tmp + 4 * d, *G0iG = tmp + 5 * d;
  double *R = G0iG + d * d, *t = R + d * d;
  for (int i = 0; i < d; ++i) n1[i] = -P->nac[i] / P->masses[i];
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += K->Gamma_0[i * d + k] * K->iGi0[k * d + j];
      G0iG[i * d + j] = s;
    }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += G0iG[i * d + k] * K->Gamma_i[k * d + j];
      R[i * d + j] = s;
    }
  for (int i = 0; i < d; ++i) {
    double s1 = 0, s2 = 0;
    for (int j = 0; j < d; ++j) { s1 += G0iG[i * d + j] * (Pm[j] - K->p0[j]); s2 += G0iG[i * d + j] * (p[j] - K->p0[j]); }
    PI[i] = K->p0[i] + s1;
    pi[i] = K->p0[i] + s2;
    dq[i] = K->q0[i] - q[i];
    dQ[i] = K->q0[i] - Q[i];
  }
  (void)t;
  double PIn = 0, pin = 0;
  for (int i = 0; i < d; ++i) { PIn += PI[i] * n1[i]; pin += pi[i] * n1[i]; }
  const cplx nacQ = quad(R, dQ, n1, d) - I * PIn;
  const cplx nacq = quad(R, dq, n1, d) + I * pin;
  *kic = nacQ * nacq * cq * w;
}

/* optional export of the pieces the WM wavefunction diagnostics need (propagators.py:1339-1342): when non-NULL, wm_prefactor
 * writes [CQQ (d^2) | CqQ (d^2) | PIq (d) | PIQ (d) | eps] of the trajectory there */
static __thread cplx *g_wm_export = NULL;

/* WM prefactor pieces for one trajectory (propagators.py:1155-1389) */
static void wm_prefactor(const sc_oracle_consts *K, traj_t *T, const double *zi, cplx *w) {
  const int d = K->d, dr = K->dr, d2 = d * d, D2 = 2 * d, R2 = 2 * dr;
  const double *Mqq = T->y + 2 * d, *Mqp = Mqq + d2, *Mpq = Mqp + d2, *Mpp = Mpq + d2;
  const double *p = zi + d, *Pm = T->y + d;
  cplx *A = w, *Ap = A + D2 * D2, *iAp = Ap + R2 * R2, *iA = iAp + R2 * R2, *U2 = iA + D2 * D2;
  cplx *BQ = U2 + D2 * R2, *Bq = BQ + d * D2, *b0 = Bq + d * D2, *t1 = b0 + D2, *t2 = t1 + D2 * D2;
  cplx *Gt = t2 + D2 * D2, *Gti = Gt + d2, *CQQ = Gti + d2, *CqQ = CQQ + d2, *Cqq = CqQ + d2;
  cplx *Mm = Cqq + d2, *Mp = Mm + d2, *iMp = Mp + dr * dr, *iM = iMp + dr * dr;
  cplx *pit = iM + d2, *pii = pit + d, *PIq = pii + d, *PIQ = PIq + d, *v1 = PIQ + d, *v2 = v1 + D2;
  cplx *G0c = v2 + D2, *iGc = G0c + d2, *GiG = iGc + d2; /* complex copies of Gamma_0, iGi0, Gamma_0 iGi0 */
#define MQZ(j, i) ((i) < d ? Mqq[(j) * d + (i)] : Mqp[(j) * d + (i) - d])
#define MPZ(j, i) ((i) < d ? Mpq[(j) * d + (i)] : Mpp[(j) * d + (i) - d])
  /* gradient and Hessian of L = i S (eqns A4-A9) */
  cplx *gradL = v1;
  for (int j = 0; j < D2; ++j) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) s += MQZ(i, j) * Pm[i];
    if (j < d) s -= p[j];
    gradL[j] = I * s;
  }
  /* A = 2 F - hessL + Mqz^T Gt Mqz + Eqz^T Gi Eqz + 2i (Mpz^T Mqz - Epz^T Eqz)   (eqn 50) */
  for (int i = 0; i < D2; ++i)
    for (int l = 0; l < D2; ++l) {
      double f = 0.0;
      if (i < d && l < d) f = K->alpha * K->Gamma_0[i * d + l];
      if (i >= d && l >= d) f = K->beta * K->iGamma_0[(i - d) * d + (l - d)];
      /* hessL[i][l] = i sum_j X[j][i] Y[j][l] with (X,Y) = (Mpq|Mqp rows, Mqz) */
      double hs = 0.0, mg = 0.0, pm = 0.0;
      for (int j = 0; j < d; ++j) {
        const double xji = (i < d) ? Mpq[j * d + i] : Mqp[j * d + i - d];
        double yjl;
        if (i < d) yjl = (l < d) ? Mqq[j * d + l] : Mqp[j * d + l - d];
        else yjl = (l < d) ? Mpq[j * d + l] : Mpp[j * d + l - d];
        hs += xji * yjl;
        pm += MPZ(j, i) * MQZ(j, l);
      }
      for (int j = 0; j < d; ++j) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += K->Gamma_t[j * d + k] * MQZ(k, l);
        mg += MQZ(j, i) * s;
      }
      double gi = (i < d && l < d) ? K->Gamma_i[i * d + l] : 0.0;
      double ee = (i >= d && l == i - d) ? 1.0 : 0.0;
      A[i * D2 + l] = 2 * f - I * hs + mg + gi + 2.0 * I * (pm - ee);
    }
  /* U2 = blockdiag(U, U);  A' = U2^T A U2 */
  memset(U2, 0, sizeof(cplx) * D2 * R2);
  for (int i = 0; i < d; ++i)
    for (int a = 0; a < dr; ++a) { U2[i * R2 + a] = K->U[i * dr + a]; U2[(d + i) * R2 + dr + a] = K->U[i * dr + a]; }
  for (int a = 0; a < R2; ++a)
    for (int j = 0; j < D2; ++j) {
      cplx s = 0.0;
      for (int i = 0; i < D2; ++i) s += U2[i * R2 + a] * A[i * D2 + j];
      t1[a * D2 + j] = s;
    }
  cgemm(R2, R2, D2, t1, U2, 0, Ap);
  memcpy(t2, Ap, sizeof(cplx) * R2 * R2);
  cinv(t2, iAp, R2);
  /* iA = U2 iA' U2^T */
  cgemm(D2, R2, R2, U2, iAp, 0, t1);
  cgemm(D2, D2, R2, t1, U2, 1, iA);
  /* BQ = Gt Mqz + i Mpz ; Bq = [Gi, -i 1] ; b0 (eqns 53-55) */
  for (int i = 0; i < d; ++i)
    for (int k = 0; k < D2; ++k) {
      double s = 0.0;
      for (int j = 0; j < d; ++j) s += K->Gamma_t[i * d + j] * MQZ(j, k);
      BQ[i * D2 + k] = s + I * MPZ(i, k);
      Bq[i * D2 + k] = (k < d) ? K->Gamma_i[i * d + k] : ((k - d == i) ? -I : 0.0);
    }
  for (int i = 0; i < D2; ++i) {
    double s = 0.0;
    for (int j = 0; j < d; ++j) s += MQZ(j, i) * Pm[j];
    if (i < d) s -= p[i];
    b0[i] = gradL[i] - I * s;
  }
  /* Gt_ = Gamma_t - BQ iA BQ^T ; Gti = BQ iA Bq^T   (eqns 57, 59) */
  cgemm(d, D2, D2, BQ, iA, 0, t1);  /* t1 = BQ iA  (d x 2d) */
  cgemm(d, d, D2, t1, BQ, 1, Gt);
  for (int i = 0; i < d2; ++i) Gt[i] = K->Gamma_t[i] - Gt[i];
  cgemm(d, d, D2, t1, Bq, 1, Gti);
  /* pi_t = P - i BQ iA b0 ; pi_i = p + i Bq iA b0  (eqn 60) */
  for (int i = 0; i < d; ++i) {
    cplx s = 0.0;
    for (int k = 0; k < D2; ++k) s += t1[i * D2 + k] * b0[k];
    pit[i] = Pm[i] - I * s;
  }
  cgemm(d, D2, D2, Bq, iA, 0, t2);
  for (int i = 0; i < d; ++i) {
    cplx s = 0.0;
    for (int k = 0; k < D2; ++k) s += t2[i * D2 + k] * b0[k];
    pii[i] = p[i] + I * s;
  }
  for (int i = 0; i < d2; ++i) { G0c[i] = K->Gamma_0[i]; iGc[i] = K->iGi0[i]; }
  cgemm(d, d, d, G0c, iGc, 0, GiG);                      /* Gamma_0 iGi0 */
  cgemm(d, d, d, GiG, G0c, 0, Cqq);
  for (int i = 0; i < d2; ++i) Cqq[i] = G0c[i] - Cqq[i];  /* eqn 69 */
  cgemm(d, d, d, Gti, iGc, 0, t1);                         /* Gti iGi0 */
  cgemm(d, d, d, t1, Gti, 1, CQQ);
  for (int i = 0; i < d2; ++i) CQQ[i] = Gt[i] - CQQ[i];    /* eqn 70 */
  cgemm(d, d, d, GiG, Gti, 1, CqQ);                        /* eqn 71 */
  for (int i = 0; i < d; ++i) v2[i] = K->p0[i] - pii[i];
  for (int i = 0; i < d; ++i) {
    cplx s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < d; ++k) { s1 += GiG[i * d + k] * v2[k]; s2 += t1[i * d + k] * v2[k]; }
    PIq[i] = K->p0[i] - s1;   /* eqn 72 */
    PIQ[i] = pit[i] + s2;     /* eqn 73 */
  }
  const cplx eps = 0.5 * cquad(iA, b0, b0, D2) - 0.5 * cquad(iGc, v2, v2, d); /* eqn 74 */
  if (g_wm_export) {
    memcpy(g_wm_export, CQQ, sizeof(cplx) * d2);
    memcpy(g_wm_export + d2, CqQ, sizeof(cplx) * d2);
    memcpy(g_wm_export + 2 * d2, PIq, sizeof(cplx) * d);
    memcpy(g_wm_export + 2 * d2 + d, PIQ, sizeof(cplx) * d);
    g_wm_export[2 * d2 + 2 * d] = eps;
  }
  /* det(A' / (2 sqrt(alpha beta))) */
  const double sc = 2.0 * sqrt(K->alpha * K->beta);
  for (int i = 0; i < R2 * R2; ++i) t2[i] = Ap[i] / sc;
  T->detA = cdet(t2, R2);
  track_sign(T, 1, T->detA);
  /* M = Gamma_0 + CQQ, projected; inverse; det(M/(2 pi))  (eqn 78) */
  for (int i = 0; i < d2; ++i) Mm[i] = G0c[i] + CQQ[i];
  for (int a = 0; a < dr; ++a)
    for (int j = 0; j < d; ++j) {
      cplx s = 0.0;
      for (int i = 0; i < d; ++i) s += K->U[i * dr + a] * Mm[i * d + j];
      t1[a * d + j] = s;
    }
  cgemm(dr, dr, d, t1, K->U, 0, Mp);
  memcpy(t2, Mp, sizeof(cplx) * dr * dr);
  cinv(t2, iMp, dr);
  for (int i = 0; i < dr * dr; ++i) t2[i] = Mp[i] / (2 * M_PI);
  T->detM = cdet(t2, dr);
  cgemm(d, dr, dr, K->U, iMp, 0, t1);
  cgemm(d, d, dr, t1, K->U, 1, iM);
  /* eqns 79-84 */
  cgemm(d, d, d, CqQ, iM, 0, t1);              /* CqQ iM */
  cgemm(d, d, d, t1, CqQ, 1, T->Rqq);
  for (int i = 0; i < d2; ++i) T->Rqq[i] = Cqq[i] - T->Rqq[i];
  cgemm(d, d, d, t1, G0c, 0, T->RqQ);
  cgemm(d, d, d, G0c, iM, 0, t2);              /* Gamma_0 iM */
  cgemm(d, d, d, t2, G0c, 0, T->RQQ);
  for (int i = 0; i < d2; ++i) T->RQQ[i] = G0c[i] - T->RQQ[i];
  for (int i = 0; i < d; ++i) v2[i] = PIQ[i] - K->p0[i];
  for (int i = 0; i < d; ++i) {
    cplx s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < d; ++k) { s1 += t1[i * d + k] * v2[k]; s2 += t2[i * d + k] * v2[k]; }
    T->Pq[i] = PIq[i] - s1;
    T->PQ[i] = K->p0[i] + s2;
  }
  T->gamma = eps - 0.5 * cquad(iM, v2, v2, d);
  track_sign(T, 2, T->detM);
#undef MQZ
#undef MPZ
}

/* WM per-trajectory contributions (propagators.py:1577-1614, 1673-1717) */
static void wm_contrib(const sc_oracle_potential *P, const sc_oracle_consts *K, const traj_t *T, const double *zi,
                       double probi, cplx *cauto, cplx *kic, cplx *tmp) {
  const int d = K->d;
  const double *q = zi, *Q = T->y;
  const double S = T->y[2 * d + 4 * d * d];
  cplx *dq = tmp, *dQ = tmp + d, *n1 = tmp + 2 * d;
  for (int i = 0; i < d; ++i) { dq[i] = K->q0[i] - q[i]; dQ[i] = K->q0[i] - Q[i]; n1[i] = -P->nac[i] / P->masses[i]; }
  cplx pref = sqrt(K->detG0) * pow(K->detGt, 0.25) * pow(K->detGi, 0.25) / sqrt(K->detGi0);
  pref *= (T->sign[0] * T->c) * cexp(I * S);
  pref *= 1.0 / csqrt(T->detA) * T->sign[1];
  pref *= 1.0 / csqrt(T->detM) * T->sign[2];
  cplx pq = 0.0, pQ = 0.0, Pqn = 0.0, PQn = 0.0;
  for (int i = 0; i < d; ++i) { pq += T->Pq[i] * dq[i]; pQ += T->PQ[i] * dQ[i]; Pqn += T->Pq[i] * n1[i]; PQn += T->PQ[i] * n1[i]; }
  const cplx expo = T->gamma - 0.5 * cquad(T->Rqq, dq, dq, d) - 0.5 * cquad(T->RQQ, dQ, dQ, d) +
                    cquad(T->RqQ, dq, dQ, d) - I * pq + I * pQ;
  const cplx cq = pref * cexp(expo);
  const double w = 1.0 / (probi * pow(2 * M_PI, d));
  *cauto = cq * w;
  const cplx nacqQ = cquad(T->RqQ, n1, n1, d);
  const cplx nacQ = cquad(T->RQQ, dQ, n1, d) - cquad(T->RqQ, dq, n1, d) - I * PQn;
  /* einsum('in,jin,jn->n', q0-Q, RqQ, n1q) = n1^T RqQ (q0-Q) */
  const cplx nacq = cquad(T->Rqq, dq, n1, d) - cquad(T->RqQ, n1, dQ, d) + I * Pqn;
  *kic = (nacqQ + nacQ * nacq) * cq * w;
}

/* ------------------------------------------------------------------ driver ----------------- */

/*
 * Runs the reference's loop  { C_auto[t], k_ic[t] = read ; step }  for nt samples (cli.py:401-436).
 *   wm          : 0 = Herman-Kluk, 1 = Walton-Manolopoulos
 *   ntraj_norm  : the N of the Monte-Carlo weight 1/(N probi (2 pi)^d) (global ensemble size when sharded)
 *   zi (2d,n), probi (n)       injected ensemble
 *   auto_out, ic_out (nt)      complex, with the e^{i t E0} phase, t accumulated by += dt
 *   energy_out (nt)            mean (T+V) of the 4th RK4 stage of the step following sample t
 *   y_out (ylen, n) or NULL, c_out/c2_out (n) or NULL, signs_out (3, n) or NULL
 * returns 0, or 1 if the energy-conservation guard (propagators.py:385-398) would have raised.
 */
int sc_oracle_run(const sc_oracle_potential *P, const sc_oracle_consts *K, int wm, int n, long ntraj_norm,
                  const double *zi, const double *probi, double dt, int nt, double energy0_es, cplx *auto_out,
                  cplx *ic_out, double *energy_out, double *y_out, cplx *c_out, cplx *c2_out, double *signs_out,
                  int nthreads) {
  const int d = K->d, L = ylen(d), d2 = d * d, D2 = 2 * d;
  int rc = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  double *Y = (double *)calloc((size_t)n * L, sizeof(double));
  double *Z = (double *)malloc(sizeof(double) * (size_t)n * 2 * d);
  traj_t *T = (traj_t *)calloc(n, sizeof(traj_t));
  cplx *wmbuf = wm ? (cplx *)calloc((size_t)n * (3 * d2 + 2 * d), sizeof(cplx)) : NULL;
  double *acc = (double *)calloc((size_t)nt * 5, sizeof(double));
  for (int t = 0; t < n; ++t) {
    double *y = Y + (size_t)t * L;
    for (int k = 0; k < 2 * d; ++k) { Z[(size_t)t * 2 * d + k] = zi[(size_t)k * n + t]; y[k] = Z[(size_t)t * 2 * d + k]; }
    for (int a = 0; a < d; ++a) { y[2 * d + a * d + a] = 1.0; y[2 * d + 3 * d2 + a * d + a] = 1.0; }
    T[t].y = y;
    if (wm) {
      cplx *b = wmbuf + (size_t)t * (3 * d2 + 2 * d);
      T[t].Rqq = b; T[t].RQQ = b + d2; T[t].RqQ = b + 2 * d2; T[t].Pq = b + 3 * d2; T[t].PQ = b + 3 * d2 + d;
    }
  }
  const size_t wsz = (size_t)16 * D2 * D2 + 64 * d + 64;
  const size_t rksz = 5 * (size_t)L + d + d2, tmpsz = 8 * (size_t)d + 3 * d2 + 8;
  int maxthr = 1;
#ifdef _OPENMP
  maxthr = omp_get_max_threads();
#endif
  cplx *W = (cplx *)malloc(sizeof(cplx) * wsz * maxthr);
  double *RK = (double *)malloc(sizeof(double) * (rksz + tmpsz) * maxthr);
#ifdef _OPENMP
#define TID omp_get_thread_num()
#else
#define TID 0
#endif
  /* t = 0: prefactor initialises the branch trackers (propagators.py:628-631) */
#pragma omp parallel for schedule(static)
  for (int t = 0; t < n; ++t) {
    cplx *w = W + wsz * TID;
    hk_prefactor(K, &T[t], w);
    if (wm) wm_prefactor(K, &T[t], Z + (size_t)t * 2 * d, w);
  }
  for (int s = 0; s < nt; ++s) {
    double a_re = 0, a_im = 0, k_re = 0, k_im = 0, en = 0;
#pragma omp parallel for schedule(static) reduction(+ : a_re, a_im, k_re, k_im, en)
    for (int t = 0; t < n; ++t) {
      cplx *w = W + wsz * TID;
      double *rk = RK + (rksz + tmpsz) * TID, *tmp = rk + rksz;
      cplx ca, ki;
      if (wm) wm_contrib(P, K, &T[t], Z + (size_t)t * 2 * d, probi[t], &ca, &ki, w);
      else hk_contrib(P, K, &T[t], Z + (size_t)t * 2 * d, probi[t], &ca, &ki, tmp);
      a_re += creal(ca); a_im += cimag(ca); k_re += creal(ki); k_im += cimag(ki);
      double e;
      rk4_step(P, T[t].y, dt, rk, &e);
      en += e;
      hk_prefactor(K, &T[t], w);
      if (wm) wm_prefactor(K, &T[t], Z + (size_t)t * 2 * d, w);
    }
    acc[5 * s + 0] = a_re; acc[5 * s + 1] = a_im; acc[5 * s + 2] = k_re; acc[5 * s + 3] = k_im; acc[5 * s + 4] = en;
  }
  free(W); free(RK);
  double tcur = 0.0;
  for (int s = 0; s < nt; ++s) {
    const cplx ph = cexp(I * tcur * energy0_es);
    auto_out[s] = (acc[5 * s + 0] + I * acc[5 * s + 1]) / (double)ntraj_norm * ph;
    ic_out[s] = (acc[5 * s + 2] + I * acc[5 * s + 3]) / (double)ntraj_norm * ph;
    if (energy_out) energy_out[s] = acc[5 * s + 4] / n;
    if (s > 0 && fabs(acc[5 * s + 4] - acc[5 * s - 1]) / n > 1.0e-2) rc = 1;
    tcur += dt;
  }
  for (int t = 0; t < n; ++t) {
    if (y_out)
      for (int k = 0; k < L; ++k) y_out[(size_t)k * n + t] = Y[(size_t)t * L + k];
    if (c_out) c_out[t] = T[t].c;
    if (c2_out) c2_out[t] = T[t].c2;
    if (signs_out)
      for (int k = 0; k < 3; ++k) signs_out[(size_t)k * n + t] = T[t].tracked[k] ? T[t].sign[k] : 1.0;
  }
  free(acc); free(wmbuf); free(T); free(Z); free(Y);
  return rc;
}

/*
 * Pieces of the Walton-Manolopoulos wavefunction diagnostics (propagators.py:1391-1575) for a given state y (ylen, n) of an
 * ensemble zi (2d, n): out (n, 2 d^2 + 2 d + 2) complex = [CQQ | CqQ | PIq | PIQ | eps | detA] per trajectory, recomputed by
 * wm_prefactor (no branch tracking: the trackers of the run are passed separately to the numpy restatement).
 */
int sc_oracle_wm_diag(const sc_oracle_consts *K, int n, const double *y_in, const double *zi, cplx *out) {
  const int d = K->d, L = ylen(d), d2 = d * d, D2 = 2 * d;
  const size_t wsz = (size_t)16 * D2 * D2 + 64 * d + 64, stride = 2 * (size_t)d2 + 2 * d + 2;
  cplx *w = (cplx *)malloc(sizeof(cplx) * wsz), *buf = (cplx *)calloc(3 * d2 + 2 * d, sizeof(cplx));
  double *y = (double *)malloc(sizeof(double) * L), *z = (double *)malloc(sizeof(double) * 2 * d);
  for (int t = 0; t < n; ++t) {
    traj_t T;
    memset(&T, 0, sizeof(T));
    for (int k = 0; k < L; ++k) y[k] = y_in[(size_t)k * n + t];
    for (int k = 0; k < 2 * d; ++k) z[k] = zi[(size_t)k * n + t];
    T.y = y;
    T.Rqq = buf; T.RQQ = buf + d2; T.RqQ = buf + 2 * d2; T.Pq = buf + 3 * d2; T.PQ = buf + 3 * d2 + d;
    hk_prefactor(K, &T, w);
    g_wm_export = out + (size_t)t * stride;
    wm_prefactor(K, &T, z, w);
    g_wm_export = NULL;
    out[(size_t)t * stride + stride - 1] = T.detA;
  }
  free(w); free(buf); free(y); free(z);
  return 0;
}

int sc_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
