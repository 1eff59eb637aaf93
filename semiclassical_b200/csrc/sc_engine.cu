// sc_engine.cu -- C ABI (include/semiclassical_b200.h) and host-side dispatch of the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/semiclassical_b200.h"
#include "sc_kernels.cuh"
#include "sc_small.cuh"
#include "sc_mma.cuh"
#include "sc_chunk.cuh"
#include "sc_stream.cuh"
#include "sc_gdml2.cuh"
#include "sc_lu_batch.cuh"
#include "sc_potentials.cuh"
#include "sc_wm.cuh"
#include "sc_gauss.cuh"

using namespace sc;

// ------------------------------------------------------------------ error handling ----------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
// like CU inside the sc_potential_create_* functions: the half-built potential `p` is released before returning
#define CUP(x)                                                                                     \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) {                                                                       \
      delete p;                                                                                    \
      return fail(SC_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    }                                                                                              \
  } while (0)
#define CU(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) return fail(SC_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

extern "C" int sc_abi_version(void) { return SC_ABI_VERSION; }
extern "C" const char *sc_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------ device buffers ----------
struct DevPool {
  std::vector<void *> ptrs;
  DevPool() = default;
  DevPool(const DevPool &) = delete;
  DevPool &operator=(const DevPool &) = delete;
  ~DevPool() { release(); }
  void release() {
    for (void *p : ptrs) cudaFree(p);
    ptrs.clear();
  }
  cudaError_t upload(const double *host, size_t n, const double **out) {
    double *p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(double) * (n ? n : 1));
    if (e != cudaSuccess) return e;
    ptrs.push_back(p);
    if (n) e = cudaMemcpy(p, host, sizeof(double) * n, cudaMemcpyHostToDevice);
    *out = p;
    return e;
  }
  template <typename T>
  cudaError_t alloc(size_t n, T **out) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(T) * (n ? n : 1));
    if (e != cudaSuccess) return e;
    ptrs.push_back(p);
    *out = reinterpret_cast<T *>(p);
    return e;
  }
};

// ------------------------------------------------------------------ Walton-Manolopoulos host state
struct WMState {
  WMDev dev;
  WMLayout L;
  double *n1_dev = nullptr, *scratch5 = nullptr, *partials = nullptr;
  int tpt = 32, groups = 1, grid = 1;
  size_t smem = 0;
  double2 *gws = nullptr;           // workspace slabs in global memory when one trajectory exceeds shared memory (k_wm_global)
  bool global_ws = false;
  // K-step fused launches: snapshots of (step, trajectory) records, per-group rows, HK rows (energies)
  void *snap = nullptr;
  size_t snap_cap = 0;
  // wavefunction diagnostics: trajectory-minor arrays written by the WM_DIAG pass + partial rows
  double *diag = nullptr;
  size_t diag_cap = 0;
};

struct sc_potential {
  PotDev dev;
  DevPool pool;
  std::vector<double> imass, n1;  // host copies for the IC constants
};

struct sc_engine {
  sc_engine_config cfg;
  EngDev dev;
  DevPool pool;       // constants
  DevPool ens;        // ensemble-sized buffers
  std::vector<double> hR, hG0iG, hp0;  // host: R = G0 iGi0 Gi, G0 iGi0, p0
  const double *oiA = nullptr, *oiB = nullptr, *oiC = nullptr;
  double *d_wR = nullptr, *d_wG = nullptr;
  const double *d_R = nullptr, *d_G0iG = nullptr;   // R = G0 iGi0 Gi and G0 iGi0 on the device (position-dependent NAC read-out)
  const sc_potential *nac_pot = nullptr;
  std::vector<double> nac_cache;
  double *partials = nullptr;
  size_t partials_cap = 0;
  void *chunk_scratch = nullptr;   // prefactor matrices, determinants, aux rows of one batch (chunked path)
  size_t chunk_scratch_cap = 0;
  // optional per-kernel timing of the chunked path (CUDA events on the launching stream)
  bool dense_engine = false;      // separable models through the general dense pipeline (sc_engine_set_option)
  bool timing = false;
  std::vector<cudaEvent_t> tev;
  size_t tev_used = 0;
  std::vector<int> tev_slot;              // slot of the kernel(s) that follow the event, -1: end of a window
  double tms[SC_TIMING_SLOTS] = {0.0};    // path (+aux), rk4, lu, finish, rmult, potential Hessians
  // dense column pipeline (sc_stream.cuh): padded constant A operands
  double *stream_const = nullptr;         // [H0 | L1 | L2], each d x ldh
  double *corr_dev = nullptr;
  size_t corr_cap = 0;
  double *diag_scratch = nullptr, *pack_scratch = nullptr;   // wavefunction diagnostics: persistent scratch
  size_t diag_cap = 0, pack_cap = 0;
  double *stage_in = nullptr;                                // sc_engine_set_ensemble_host: device staging of (zi, probi)
  size_t stage_in_cap = 0;
  long long ntraj_norm = 0;
  int ens_n = 0;                  // size of the ensemble the buffers in `ens` were allocated for
  long long launches = 0;
  int sm_count = 148;
  const char *kernel_name = "none";
  WMState wm;
  // generic-potential stage path
  double *stage_buf = nullptr;
  ~sc_engine() {
    if (partials) cudaFree(partials);
    if (corr_dev) cudaFree(corr_dev);
    if (chunk_scratch) cudaFree(chunk_scratch);
    if (stream_const) cudaFree(stream_const);
    if (wm.snap) cudaFree(wm.snap);
    if (wm.diag) cudaFree(wm.diag);
    if (diag_scratch) cudaFree(diag_scratch);
    if (pack_scratch) cudaFree(pack_scratch);
    if (stage_in) cudaFree(stage_in);
    for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
  }
};

// ------------------------------------------------------------------ potentials --------------
static int pot_common(sc_potential *p, int type, int d, const double *masses, const double *nac) {
  p->dev = PotDev();
  p->dev.type = type;
  p->dev.d = d;
  p->imass.resize(d);
  p->n1.resize(d);
  for (int i = 0; i < d; ++i) {
    p->imass[i] = 1.0 / (masses ? masses[i] : 1.0);
    p->n1[i] = -(nac ? nac[i] : 1.0) * p->imass[i];
  }
  CU(p->pool.upload(p->imass.data(), d, &p->dev.imass));
  CU(p->pool.upload(p->n1.data(), d, &p->dev.n1));
  return SC_OK;
}

static int check_dim(int d) {
  if (d < 1 || d > SC_MAX_DIM) return fail(SC_ERR_UNSUPPORTED, "dimension %d outside [1, %d]", d, SC_MAX_DIM);
  return SC_OK;
}

extern "C" int sc_potential_create_morse(sc_potential **out, int d, const double *omega, const double *a,
                                         const double *D, int all_harmonic, const double *nac) {
  if (!out || !omega || !a || !D || !nac) return fail(SC_ERR_INVALID, "null argument");
  if (int rc = check_dim(d)) return rc;
  sc_potential *p = new sc_potential();
  int rc = pot_common(p, POT_MORSE, d, nullptr, nac);
  if (rc) { delete p; return rc; }
  CUP(p->pool.upload(omega, d, &p->dev.omega));
  CUP(p->pool.upload(a, d, &p->dev.a));
  CUP(p->pool.upload(D, d, &p->dev.D));
  p->dev.all_harmonic = all_harmonic;
  *out = p;
  return SC_OK;
}

extern "C" int sc_potential_create_rotated_morse(sc_potential **out, int d, const double *omega, const double *a,
                                                 const double *D, int all_harmonic, const double *nac,
                                                 const double *Q) {
  if (!Q) return fail(SC_ERR_INVALID, "null argument");
  int rc = sc_potential_create_morse(out, d, omega, a, D, all_harmonic, nac);
  if (rc) return rc;
  sc_potential *p = *out;
  p->dev.type = POT_ROTATED_MORSE;
  *out = nullptr;
  CUP(p->pool.upload(Q, (size_t)d * d, &p->dev.Q));
  *out = p;
  return SC_OK;
}

extern "C" int sc_potential_create_nonharmonic(sc_potential **out, int d, const double *eps, const double *b) {
  if (!out || !eps || !b) return fail(SC_ERR_INVALID, "null argument");
  if (int rc = check_dim(d)) return rc;
  sc_potential *p = new sc_potential();
  int rc = pot_common(p, POT_NONHARMONIC, d, nullptr, nullptr);  // masses 1, tau1 = 1
  if (rc) { delete p; return rc; }
  CUP(p->pool.upload(eps, d, &p->dev.eps));
  CUP(p->pool.upload(b, d, &p->dev.b));
  *out = p;
  return SC_OK;
}

extern "C" int sc_potential_create_harmonic(sc_potential **out, int d, const double *pos0, double energy0,
                                            const double *grad0, const double *hess0, const double *masses,
                                            const double *nac) {
  if (!out || !pos0 || !grad0 || !hess0 || !masses || !nac) return fail(SC_ERR_INVALID, "null argument");
  if (int rc = check_dim(d)) return rc;
  sc_potential *p = new sc_potential();
  int rc = pot_common(p, POT_HARMONIC, d, masses, nac);
  if (rc) { delete p; return rc; }
  CUP(p->pool.upload(pos0, d, &p->dev.pos0));
  CUP(p->pool.upload(grad0, d, &p->dev.grad0));
  CUP(p->pool.upload(hess0, (size_t)d * d, &p->dev.hess0));
  p->dev.e0 = energy0;
  *out = p;
  return SC_OK;
}

extern "C" int sc_potential_create_gdml(sc_potential **out, int n_atoms, int n_train, int n_desc,
                                        const double *xs_train, const double *jx_alphas, double sig, double c,
                                        double std, const double *masses, const double *nac) {
  if (!out || !xs_train || !jx_alphas || !masses || !nac) return fail(SC_ERR_INVALID, "null argument");
  const int d = 3 * n_atoms;
  if (int rc = check_dim(d)) return rc;
  if (n_desc != n_atoms * (n_atoms - 1) / 2) return fail(SC_ERR_INVALID, "n_desc != N(N-1)/2");
  sc_potential *p = new sc_potential();
  int rc = pot_common(p, POT_GDML, d, masses, nac);
  if (rc) { delete p; return rc; }
  CUP(p->pool.upload(xs_train, (size_t)n_train * n_desc, &p->dev.xs_train));
  CUP(p->pool.upload(jx_alphas, (size_t)n_train * n_desc, &p->dev.jx_alphas));
  p->dev.n_atoms = n_atoms;
  p->dev.n_train = n_train;
  p->dev.n_desc = n_desc;
  p->dev.sig = sig;
  p->dev.e0 = c;
  p->dev.gstd = std;
  *out = p;
  return SC_OK;
}

extern "C" int sc_potential_set_origin(sc_potential *pot, double origin) {
  if (!pot) return fail(SC_ERR_INVALID, "null potential");
  pot->dev.origin = origin;
  return SC_OK;
}
extern "C" int sc_potential_dimensions(const sc_potential *pot) { return pot ? pot->dev.d : -1; }
extern "C" int sc_potential_destroy(sc_potential *pot) {
  delete pot;
  return SC_OK;
}

extern "C" int sc_potential_eval(const sc_potential *pot, int n, const double *r, double *V, double *grad,
                                 double *hess, void *stream) {
  if (!pot || !r || !V) return fail(SC_ERR_INVALID, "null argument");
  if (n <= 0) return SC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = launch_potential_eval(pot->dev, n, r, V, grad, hess, st);
  if (rc) return fail(SC_ERR_UNSUPPORTED, "potential type %d not supported by sc_potential_eval", pot->dev.type);
  CU(cudaGetLastError());
  return SC_OK;
}


// ------------------------------------------------------------------ Walton-Manolopoulos dispatch
// constants of propagators.py:1102-1130 and the trajectory-independent matrices of eqns (68, 69)
static int wm_setup(WMState &w, const sc_engine_config &cfg, DevPool &pool) {
  const int d = cfg.d, dr = cfg.dr;
  const size_t dd = (size_t)d * d;
  if (!cfg.iGamma_0 || !cfg.U) return fail(SC_ERR_INVALID, "WM needs iGamma_0 and U");
  w.dev = WMDev();
  w.dev.d = d;
  w.dev.dr = dr;
  w.dev.alpha = cfg.alpha;
  w.dev.beta = cfg.beta;
  w.dev.pref = std::sqrt(cfg.detG0) * std::pow(cfg.detGt, 0.25) * std::pow(cfg.detGi, 0.25) / std::sqrt(cfg.detGi0);
  w.dev.pref_coef = std::pow(cfg.detG0, 0.25) * std::pow(cfg.detGt, 0.25) * std::pow(cfg.detGi, 0.25) / std::sqrt(cfg.detGi0);
  std::vector<double> GiG(dd, 0.0), Cqq(dd, 0.0);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += cfg.Gamma_0[i * d + k] * cfg.iGi0[k * d + j];
      GiG[i * d + j] = s;
    }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += GiG[i * d + k] * cfg.Gamma_0[k * d + j];
      Cqq[i * d + j] = cfg.Gamma_0[i * d + j] - s;
    }
  CU(pool.upload(cfg.Gamma_0, dd, &w.dev.G0));
  CU(pool.upload(cfg.Gamma_i, dd, &w.dev.Gi));
  CU(pool.upload(cfg.Gamma_t, dd, &w.dev.Gt));
  CU(pool.upload(cfg.iGi0, dd, &w.dev.iGi0));
  CU(pool.upload(cfg.iGamma_0, dd, &w.dev.iG0));
  CU(pool.upload(GiG.data(), dd, &w.dev.GiG));
  CU(pool.upload(Cqq.data(), dd, &w.dev.Cqq));
  CU(pool.upload(cfg.U, (size_t)d * dr, &w.dev.U));
  CU(pool.alloc((size_t)d, &w.n1_dev));
  CU(cudaMemset(w.n1_dev, 0, sizeof(double) * d));
  w.dev.n1 = w.n1_dev;
  w.L = make_wm_layout(d, dr);
  const size_t ws = sizeof(double2) * (size_t)w.L.total;
  if (d <= 8) {
    w.tpt = 32;
    w.groups = (int)((200 * 1024) / ws);
    if (w.groups > 4) w.groups = 4;
    if (w.groups < 1) w.groups = 1;
  } else {
    // one CTA per trajectory; above 16 modes the workspace leaves room for one CTA per SM only: 512 threads hide the latency
    // of the shared-memory products that 128 cannot (16 of them per SM instead of 4 warps)
    w.tpt = d <= 16 ? 128 : 512;
    if (const char *s = getenv("SC_WM_TPT")) w.tpt = (atoi(s) == 512 || atoi(s) == 256) && d > 16 ? atoi(s) : 128;
    w.groups = 1;
  }
  w.smem = ws * w.groups;
  w.global_ws = false;
  if (w.smem > 227 * 1024) {
    // beyond about 29 modes (21 for rank-deficient widths) the workspace of one trajectory does not fit in shared memory: slabs in global memory
    w.global_ws = true;
    w.tpt = 256;
    w.groups = 1;
    w.smem = 0;
  }
  return SC_OK;
}

static int wm_set_nac(WMState &w, const double *n1, int d) {
  CU(cudaMemcpy(w.n1_dev, n1, sizeof(double) * d, cudaMemcpyHostToDevice));
  return SC_OK;
}

static int wm_alloc(WMState &w, DevPool &ens, const EngDev &D, const double *q0_dev, const double *p0_dev,
                    const double *probi_dev, int sm_count, cudaStream_t st) {
  const int n = D.n;
  double *winv = nullptr;
  CU(ens.alloc((size_t)n, &w.dev.prevA));
  CU(ens.alloc((size_t)n, &w.dev.prevM));
  CU(ens.alloc((size_t)n, &w.dev.signA));
  CU(ens.alloc((size_t)n, &w.dev.signM));
  CU(ens.alloc((size_t)n, &winv));
  CU(ens.alloc((size_t)8, &w.scratch5));
  w.dev.winv = winv;
  w.dev.q0 = q0_dev;
  w.dev.p0 = p0_dev;
  int per_sm = (int)((227 * 1024) / (w.smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2048 / (w.tpt * w.groups)) per_sm = 2048 / (w.tpt * w.groups);
  if (w.global_ws) {
    per_sm = 3;
    if (const char *s = getenv("SC_WM_CTAS")) per_sm = std::max(1, std::min(8, atoi(s)));
  }
  int grid = sm_count * per_sm;
  const int need = (n + w.groups - 1) / w.groups;
  if (grid > need) grid = need;
  w.grid = grid < 1 ? 1 : grid;
  if (w.global_ws) {
    double *slabs = nullptr;
    CU(ens.alloc(2 * (size_t)w.L.total * w.grid, &slabs));
    w.gws = reinterpret_cast<double2 *>(slabs);
  }
  CU(ens.alloc((size_t)w.grid * w.groups * 4, &w.partials));
  k_wm_winv<<<(n + 255) / 256, 256, 0, st>>>(probi_dev, std::pow(2.0 * M_PI, -(double)D.d), n, winv);
  CU(cudaGetLastError());
  return SC_OK;
}

// mode WM_INIT: prefactor pieces + tracker initialisation; WM_STEP: + tracker update + contributions;
// WM_CORR: contributions only (trackers untouched).  out5: device, [0..3] correlation sums, [4] <- energy_src[4]
static int wm_launch(WMState &w, const EngDev &D, int mode, double inv_norm, double *out5, const double *energy_src,
                     cudaStream_t st) {
  const int threads = w.tpt * w.groups;
  if (w.global_ws) {
    k_wm_global<256><<<w.grid, 256, 0, st>>>(D, w.dev, w.L, mode, w.partials, w.gws);
  } else if (w.tpt == 32) {
    CU(cudaFuncSetAttribute(k_wm<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
    k_wm<32><<<w.grid, threads, w.smem, st>>>(D, w.dev, w.L, mode, w.partials);
  } else if (w.tpt == 512) {
    CU(cudaFuncSetAttribute(k_wm<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
    k_wm<512><<<w.grid, threads, w.smem, st>>>(D, w.dev, w.L, mode, w.partials);
  } else if (w.tpt == 256) {
    CU(cudaFuncSetAttribute(k_wm<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
    k_wm<256><<<w.grid, threads, w.smem, st>>>(D, w.dev, w.L, mode, w.partials);
  } else {
    CU(cudaFuncSetAttribute(k_wm<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
    k_wm<128><<<w.grid, threads, w.smem, st>>>(D, w.dev, w.L, mode, w.partials);
  }
  CU(cudaGetLastError());
  if (mode != WM_INIT) {
    k_wm_reduce<<<1, 160, 0, st>>>(w.partials, w.grid * w.groups, inv_norm, energy_src, out5);
    CU(cudaGetLastError());
  }
  return SC_OK;
}

// ------------------------------------------------------------------ engine ------------------
static bool is_diagonal(const double *A, int d) {
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j)
      if (i != j && A[i * d + j] != 0.0) return false;
  return true;
}

extern "C" int sc_engine_create(sc_engine **out, const sc_engine_config *cfg) {
  if (!out || !cfg) return fail(SC_ERR_INVALID, "null argument");
  const int d = cfg->d, dr = cfg->dr;
  if (int rc = check_dim(d)) return rc;
  if (dr < 1 || dr > d) return fail(SC_ERR_INVALID, "rank %d outside [1, %d]", dr, d);
  sc_engine *e = new sc_engine();
  e->cfg = *cfg;
  EngDev &D = e->dev;
  D = EngDev();
  D.d = d;
  D.dr = dr;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, dev);
  // diagonal fast path: all three width matrices diagonal and of full rank (AS models: Gamma = diag(omega))
  const bool diag = dr == d && is_diagonal(cfg->Gamma_0, d) && is_diagonal(cfg->Gamma_i, d) && is_diagonal(cfg->Gamma_t, d);
  D.diag = diag ? 1 : 0;
  const size_t dd = (size_t)d * d;
#define UP(ptr, src, n)                                             \
  do {                                                              \
    cudaError_t ce = e->pool.upload(src, n, &ptr);                  \
    if (ce != cudaSuccess) { delete e; return fail(SC_ERR_CUDA, "upload failed: %s", cudaGetErrorString(ce)); } \
  } while (0)
  UP(D.L1, cfg->L1, (size_t)dr * d);
  UP(D.L2, cfg->L2, (size_t)dr * d);
  UP(D.R1, cfg->R1, (size_t)d * dr);
  UP(D.R2, cfg->R2, (size_t)d * dr);
  UP(D.q0, cfg->q0, d);
  UP(D.p0, cfg->p0, d);
  if (diag) {
    std::vector<double> sgt(d), isgt(d), sgi(d), isgi(d), a(d), b(d), c(d);
    for (int i = 0; i < d; ++i) {
      sgt[i] = std::sqrt(cfg->Gamma_t[i * d + i]);
      isgt[i] = 1.0 / sgt[i];
      sgi[i] = std::sqrt(cfg->Gamma_i[i * d + i]);
      isgi[i] = 1.0 / sgi[i];
    }
    UP(D.sgt, sgt.data(), d); UP(D.isgt, isgt.data(), d); UP(D.sgi, sgi.data(), d); UP(D.isgi, isgi.data(), d);
    for (int i = 0; i < d; ++i) { a[i] = cfg->ot0_A[i * d + i]; b[i] = cfg->ot0_B[i * d + i]; c[i] = cfg->ot0_C[i * d + i]; }
    UP(D.otA, a.data(), d); UP(D.otB, b.data(), d); UP(D.otC, c.data(), d);
    for (int i = 0; i < d; ++i) { a[i] = cfg->oi0_A[i * d + i]; b[i] = cfg->oi0_B[i * d + i]; c[i] = cfg->oi0_C[i * d + i]; }
    UP(e->oiA, a.data(), d); UP(e->oiB, b.data(), d); UP(e->oiC, c.data(), d);
  } else {
    UP(D.otA, cfg->ot0_A, dd); UP(D.otB, cfg->ot0_B, dd); UP(D.otC, cfg->ot0_C, dd);
    UP(e->oiA, cfg->oi0_A, dd); UP(e->oiB, cfg->oi0_B, dd); UP(e->oiC, cfg->oi0_C, dd);
  }
  D.ot_fac = cfg->ot0_fac;
  // host copies for the NAC-dependent vectors wR = R n1, wG = (G0 iGi0)^T n1  (propagators.py:894-903)
  e->hG0iG.assign(dd, 0.0);
  e->hR.assign(dd, 0.0);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += cfg->Gamma_0[i * d + k] * cfg->iGi0[k * d + j];
      e->hG0iG[i * d + j] = s;
    }
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double s = 0.0;
      for (int k = 0; k < d; ++k) s += e->hG0iG[i * d + k] * cfg->Gamma_i[k * d + j];
      e->hR[i * d + j] = s;
    }
  e->hp0.assign(cfg->p0, cfg->p0 + d);
  UP(e->d_R, e->hR.data(), dd);
  UP(e->d_G0iG, e->hG0iG.data(), dd);
  {
    cudaError_t ce = e->pool.alloc((size_t)d, &e->d_wR);
    if (ce == cudaSuccess) ce = e->pool.alloc((size_t)d, &e->d_wG);
    if (ce != cudaSuccess) { delete e; return fail(SC_ERR_CUDA, "alloc failed: %s", cudaGetErrorString(ce)); }
  }
  D.wR = e->d_wR;
  D.wG = e->d_wG;
  if (cfg->wm) {
    int rc = wm_setup(e->wm, *cfg, e->pool);
    if (rc) { delete e; return rc; }
  }
#undef UP
  // the config's pointers are the caller's; never dereference them after create
  e->cfg.L1 = e->cfg.L2 = e->cfg.R1 = e->cfg.R2 = e->cfg.U = e->cfg.q0 = e->cfg.p0 = nullptr;
  e->cfg.oi0_A = e->cfg.oi0_B = e->cfg.oi0_C = e->cfg.ot0_A = e->cfg.ot0_B = e->cfg.ot0_C = nullptr;
  e->cfg.Gamma_0 = e->cfg.Gamma_i = e->cfg.Gamma_t = e->cfg.iGi0 = e->cfg.iGamma_0 = nullptr;
  *out = e;
  return SC_OK;
}

extern "C" int sc_engine_destroy(sc_engine *e) {
  delete e;
  return SC_OK;
}

extern "C" int sc_engine_num_trajectories(const sc_engine *e) { return e ? e->dev.n : -1; }
extern "C" long long sc_engine_launch_count(const sc_engine *e) { return e ? e->launches : -1; }
extern "C" const char *sc_engine_kernel_name(const sc_engine *e) { return e ? e->kernel_name : ""; }

// NAC-dependent constants for the potential in use
static int set_nac(sc_engine *e, const double *n1, cudaStream_t st) {
  const int d = e->dev.d;
  if ((int)e->nac_cache.size() == d && std::memcmp(e->nac_cache.data(), n1, sizeof(double) * d) == 0) return SC_OK;
  std::vector<double> w(2 * d, 0.0);
  double p0n1 = 0.0;
  for (int i = 0; i < d; ++i) {
    double s = 0.0, g = 0.0;
    for (int j = 0; j < d; ++j) { s += e->hR[i * d + j] * n1[j]; g += e->hG0iG[j * d + i] * n1[j]; }
    w[i] = s;
    w[d + i] = g;
    p0n1 += e->hp0[i] * n1[i];
  }
  // synchronous copies: the previous values may still be in use by kernels in flight on `st`
  CU(cudaStreamSynchronize(st));
  CU(cudaMemcpy(e->d_wR, w.data(), sizeof(double) * d, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->d_wG, w.data() + d, sizeof(double) * d, cudaMemcpyHostToDevice));
  e->dev.p0n1 = p0n1;
  e->nac_cache.assign(n1, n1 + d);
  if (e->cfg.wm)
    if (int rc = wm_set_nac(e->wm, n1, d)) return rc;
  return SC_OK;
}

// per-kernel timing (sc_engine_set_timing): an event before every kernel group, tagged with the slot its time goes to
enum { TS_PATH = 0, TS_RK4 = 1, TS_LU = 2, TS_FINISH = 3, TS_RMULT = 4, TS_POT = 5, TS_END = -1 };
static void timing_mark(sc_engine *e, int slot, cudaStream_t st) {
  if (!e->timing) return;
  if (e->tev_used == e->tev.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    e->tev.push_back(ev);
    e->tev_slot.push_back(TS_END);
  }
  e->tev_slot[e->tev_used] = slot;
  cudaEventRecord(e->tev[e->tev_used++], st);
}

struct LaunchPlan {
  int tpt, ept, groups_per_cta, threads, grid;
  size_t smem;
  SmemLayout L;
};

static int plan_launch(const sc_engine *e, LaunchPlan &pl) {
  const int d = e->dev.d, dr = e->dev.dr, n = e->dev.n;
  const int ne = 2 * d * d;
  const int ldu = 2 * d, ldh = d;
  if (d <= 16) {
    pl.tpt = 32; pl.groups_per_cta = 4; pl.threads = 128;
    pl.ept = (ne + 31) / 32;
  } else {
    pl.tpt = (d <= 45) ? 256 : 320;
    pl.groups_per_cta = 1; pl.threads = pl.tpt;
    pl.ept = (ne + pl.tpt - 1) / pl.tpt;
  }
  pl.L = make_layout(d, dr, ldu, ldh, 0);
  pl.smem = sizeof(double) * (size_t)pl.L.total * pl.groups_per_cta;
  if (pl.smem > 227 * 1024) return fail(SC_ERR_UNSUPPORTED, "shared-memory footprint %zu B exceeds 227 KB (d = %d)", pl.smem, d);
  const int groups_needed = n;
  int ctas_per_sm = (int)((227 * 1024) / (pl.smem + 1024));
  const int max_by_threads = 2048 / pl.threads;
  if (ctas_per_sm > max_by_threads) ctas_per_sm = max_by_threads;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 8) ctas_per_sm = 8;
  int grid = e->sm_count * ctas_per_sm;
  const int need = (groups_needed + pl.groups_per_cta - 1) / pl.groups_per_cta;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  pl.grid = grid;
  return SC_OK;
}

template <int TPT, int EPT>
static cudaError_t launch_generic(const LaunchPlan &pl, const EngDev &E, const PotDev &P, double h, int nsteps, int mode,
                                  double *partials, cudaStream_t st) {
  auto kern = k_hk_generic<TPT, EPT>;
  cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (ce != cudaSuccess) return ce;
  kern<<<pl.grid, pl.threads, pl.smem, st>>>(E, P, h, nsteps, mode, partials, pl.L);
  return cudaGetLastError();
}

static int ensure_partials(sc_engine *e, size_t need, cudaStream_t st) {
  if (need > e->partials_cap) {
    CU(cudaStreamSynchronize(st));
    if (e->partials) cudaFree(e->partials);
    e->partials = nullptr;
    CU(cudaMalloc(&e->partials, sizeof(double) * need));
    e->partials_cap = need;
  }
  return SC_OK;
}

// column pipeline (sc_chunk.cuh): (q,p) path kernel -> RK4/monodromy kernel -> batched LU -> branch tracking +
// contributions, window by window over the ensemble, KC time steps per pass.  (Running the LU of window i next to
// the RK4 kernel of window i + 1 on a second stream was measured 20 % slower: the LU's pivot chain shares the FP64
// pipe with the DMMA stream.)
static int run_hk_chunked(sc_engine *e, const PotDev &P, double h, int nsteps, double *out_dev, cudaStream_t st) {
  const int d = e->dev.d, n = e->dev.n, sm = e->sm_count;
  const WColsLayout LW = make_wcols_layout(d, 1);       // one column tile per warp (two: measured 8 % slower, 8 warps / SM)
  // time steps per pass over the state: the records are read and written once per pass, so longer passes amortise the
  // state traffic and the pipeline fill (K = 8 -> 10 -> 20: +2.3 %, +4 %); passes of a launch are balanced
  int KC = 20;
  if (const char *s = getenv("SC_CHUNK_K")) KC = atoi(s) > 0 ? atoi(s) : KC;
  {
    const int npass = (nsteps + KC - 1) / KC;
    KC = (nsteps + npass - 1) / npass;
  }
  const int dp = (d + 1) & ~1;
  const size_t per_traj = (size_t)KC * ((size_t)d * d * sizeof(double2) + sizeof(double2) + 8 * sizeof(double) + 4 * dp * sizeof(double));
  size_t budget = (size_t)6 << 30   /* windows of >= 5 000 trajectories at 20 steps per pass (60 modes) */;
  if (const char *s = getenv("SC_CHUNK_SCRATCH_MB")) budget = (size_t)atol(s) << 20;
  long long ntb = (long long)(budget / per_traj);
  ntb = (ntb / sm) * sm;
  if (ntb < sm) ntb = sm;
  if (ntb > n) ntb = n;
  const size_t need_bytes = per_traj * (size_t)ntb + 512;
  if (need_bytes > e->chunk_scratch_cap) {
    CU(cudaStreamSynchronize(st));
    if (e->chunk_scratch) cudaFree(e->chunk_scratch);
    e->chunk_scratch = nullptr;
    // sized for the largest step count per pass so that a later call with another K does not reallocate
    const size_t per20 = per_traj / KC * 20;
    size_t want = std::max(need_bytes, std::min(budget, per20 * (size_t)n) + 512);
    CU(cudaMalloc(&e->chunk_scratch, want));
    e->chunk_scratch_cap = want;
  }
  double2 *cm = reinterpret_cast<double2 *>(e->chunk_scratch);
  double2 *det = cm + (size_t)KC * ntb * d * d;
  double *aux = reinterpret_cast<double *>(det + (size_t)KC * ntb);
  double *hd = aux + (size_t)KC * ntb * 8;
  unsigned long long *queue = reinterpret_cast<unsigned long long *>(hd + (size_t)KC * ntb * 4 * dp);   // work queue of k_rk4_wcols
  size_t ngroups = 0;
  for (long long t0 = 0; t0 < n; t0 += ntb) ngroups += (size_t)((std::min<long long>(ntb, n - t0) + 127) / 128);
  if (int rc = ensure_partials(e, ngroups * nsteps * 5, st)) return rc;
  const long long rk4_per_sm = 3, lu_per_sm = 0;        // CTAs per SM of k_rk4_wcols; 0: the LU launcher's default occupancy
  for (int s0 = 0; s0 < nsteps; s0 += KC) {
    const int ks = std::min(KC, nsteps - s0);
    size_t g0 = 0;
    for (long long t0 = 0; t0 < n; t0 += ntb) {
      const int nt = (int)std::min<long long>(ntb, n - t0);
      long long grid = ((long long)nt * LW.nitem + 3) / 4;
      if (grid > rk4_per_sm * sm) grid = rk4_per_sm * sm;
      timing_mark(e, TS_PATH, st);
      k_qp_path<<<(nt + 3) / 4, 128, 0, st>>>(e->dev, P, h, ks, (int)t0, nt, hd, aux);
      CU(cudaGetLastError());
      timing_mark(e, TS_RK4, st);
      CU(launch_wcols((int)grid, e->dev, P, h, ks, (int)t0, nt, cm, hd, LW, queue, st));
      timing_mark(e, TS_LU, st);
      CU(launch_lu_batch(cm, d, ks * nt, det, sm, (int)lu_per_sm, st));
      timing_mark(e, TS_FINISH, st);
      const int nblk = (nt + 127) / 128;
      k_hk_finish<<<nblk, 128, 0, st>>>(e->dev, (int)t0, nt, ks, s0, nsteps, det, aux, e->partials + g0 * nsteps * 5);
      CU(cudaGetLastError());
      timing_mark(e, TS_END, st);
      g0 += nblk;
      e->launches += 4;
    }
  }
  k_reduce_partials<<<nsteps, 160, 0, st>>>(e->partials, (int)ngroups, nsteps, 1.0 / (double)e->ntraj_norm, 1.0 / (double)n, out_dev);
  CU(cudaGetLastError());
  e->launches += 1;
  e->kernel_name = d > 24 ? "k_rk4_wcols+k_lu_mma+k_hk_finish" : "k_rk4_wcols+k_lu_warp+k_hk_finish";
  return SC_OK;
}

// dense column pipeline (sc_stream.cuh): path kernel -> overlap terms -> k_rk4_stream -> (dense widths: k_rmult) -> batched
// LU -> branch tracking + contributions, window by window, KC time steps per pass
__global__ void k_pad_matrix(const double *__restrict__ src, int rows, int cols, double *__restrict__ dst, int drows, int ld) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < drows * ld; i += gridDim.x * blockDim.x) {
    const int r = i / ld, c = i - r * ld;
    dst[i] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.0;
  }
}

// shared-memory plan of k_rk4_stream and the padded constant A operands [H0 | L1 | L2] (each d x ldh)
static int stream_setup(sc_engine *e, StreamLayout &L, cudaStream_t st) {
  const int d = e->dev.d, dr = e->dev.dr;
  int ns = 3;
  if (const char *s = getenv("SC_STREAM_SLOTS")) ns = std::max(2, std::min(4, atoi(s)));
  if (d <= 64) {
    L = make_stream_layout(d, dr, 2, ns);
    while (L.ns > 2 && sizeof(double) * (size_t)L.total > 226 * 1024) L = make_stream_layout(d, dr, 2, L.ns - 1);
  } else {
    // one tile per warp, as many warps (<= 8) as fit next to a double-buffered Hessian ring; several CTAs per trajectory
    int nw = 8;
    L = make_stream_layout(d, dr, 1, 2, nw);
    while (nw > 1 && sizeof(double) * (size_t)L.total > 226 * 1024) L = make_stream_layout(d, dr, 1, 2, --nw);
  }
  if (sizeof(double) * (size_t)L.total > 227 * 1024)
    return fail(SC_ERR_UNSUPPORTED, "stream pipeline: %zu B of shared memory (d = %d)", sizeof(double) * (size_t)L.total, d);
  const size_t hsz = (size_t)L.hsz;
  if (!e->stream_const) {
    CU(cudaMalloc(&e->stream_const, sizeof(double) * 3 * hsz));
    if (!e->dev.diag) {
      k_pad_matrix<<<32, 256, 0, st>>>(e->dev.L1, dr, d, e->stream_const + hsz, d, L.ldh);
      k_pad_matrix<<<32, 256, 0, st>>>(e->dev.L2, dr, d, e->stream_const + 2 * hsz, d, L.ldh);
      CU(cudaGetLastError());
    }
  }
  return SC_OK;
}

static int ensure_chunk_scratch(sc_engine *e, size_t need_bytes, cudaStream_t st) {
  if (need_bytes > e->chunk_scratch_cap) {
    CU(cudaStreamSynchronize(st));
    if (e->chunk_scratch) cudaFree(e->chunk_scratch);
    e->chunk_scratch = nullptr;
    e->chunk_scratch_cap = 0;
    CU(cudaMalloc(&e->chunk_scratch, need_bytes));
    e->chunk_scratch_cap = need_bytes;
  }
  return SC_OK;
}

// prefactor (+ branch tracking) of the records as they are, for d the generic kernel's MODE_INIT / MODE_TRACK cannot hold in
// shared memory: k_rk4_stream in its prefactor-only mode -> (k_rmult) -> batched LU -> k_track_only
static int run_prefactor_stream(sc_engine *e, int mode, cudaStream_t st) {
  const int d = e->dev.d, dr = e->dev.dr, n = e->dev.n, sm = e->sm_count;
  const bool dense = !e->dev.diag;
  StreamLayout L;
  if (int rc = stream_setup(e, L, st)) return rc;
  const size_t hsz = (size_t)L.hsz;
  const size_t tsz = dense ? (size_t)L.mtr * L.nt * 128 : 0, cmsz = (size_t)dr * dr * 2;
  const size_t per_traj = sizeof(double) * (cmsz + 2 + tsz);
  long long ntb = (long long)(((size_t)2 << 30) / per_traj);
  if (ntb > n) ntb = n;
  if (int rc = ensure_chunk_scratch(e, per_traj * (size_t)ntb + 1024, st)) return rc;
  double *base = reinterpret_cast<double *>(e->chunk_scratch);
  double2 *cm = reinterpret_cast<double2 *>(base);            base += (size_t)ntb * cmsz;
  double2 *det = reinterpret_cast<double2 *>(base);           base += (size_t)ntb * 2;
  StreamArgs A;
  A.hs = e->stream_const;
  A.hs_const = 1;
  A.L1p = dense ? e->stream_const + hsz : nullptr;
  A.L2p = dense ? e->stream_const + 2 * hsz : nullptr;
  A.cm = cm;
  A.T = dense ? base : nullptr;
  A.skip_rk4 = 1;
  PotDev none = PotDev();
  none.d = d;
  none.imass = e->dev.q0;
  for (long long t0 = 0; t0 < n; t0 += ntb) {
    const int nt = (int)std::min<long long>(ntb, n - t0);
    CU(launch_stream((long long)nt * L.ngroups, sm, e->dev, none, 0.0, 1, (int)t0, nt, A, L, st));
    if (dense) CU(launch_rmult(e->dev, nt, A.T, cm, sm, st));
    CU(launch_lu_batch(cm, dr, nt, det, sm, 0, st));
    k_track_only<<<(nt + 127) / 128, 128, 0, st>>>(e->dev, (int)t0, nt, det, mode == MODE_INIT ? 1 : 0);
    CU(cudaGetLastError());
    e->launches += dense ? 4 : 3;
  }
  e->kernel_name = "k_rk4_stream(prefactor)+k_lu+k_track_only";
  return SC_OK;
}

// MODE_CORR for d the generic kernel cannot hold: contributions of the current state
static int run_corr_now(sc_engine *e, double *out_dev, cudaStream_t st) {
  const int n = e->dev.n;
  int grid = std::min((n + 7) / 8, e->sm_count * 8);
  if (grid < 1) grid = 1;
  if (int rc = ensure_partials(e, (size_t)grid * 5, st)) return rc;
  k_corr_now<<<grid, 256, 0, st>>>(e->dev, e->partials);
  CU(cudaGetLastError());
  k_reduce_partials<<<1, 160, 0, st>>>(e->partials, grid, 1, 1.0 / (double)e->ntraj_norm, 1.0 / (double)n, out_dev);
  CU(cudaGetLastError());
  e->launches += 2;
  return SC_OK;
}

static int run_hk_stream(sc_engine *e, const PotDev &P, double h, int nsteps, double *out_dev, cudaStream_t st) {
  const int d = e->dev.d, dr = e->dev.dr, n = e->dev.n, sm = e->sm_count;
  const bool dense = !e->dev.diag;
  const bool hconst = P.type == POT_HARMONIC;
  StreamLayout L;
  if (int rc = stream_setup(e, L, st)) return rc;
  const size_t hsz = (size_t)L.hsz;
  if (hconst) {
    k_pad_matrix<<<32, 256, 0, st>>>(P.hess0, d, d, e->stream_const, d, L.ldh);
    CU(cudaGetLastError());
  }
  int KC = hconst ? 16 : 8;
  if (const char *s = getenv("SC_CHUNK_K")) KC = atoi(s) > 0 ? atoi(s) : KC;
  {
    const int npass = (nsteps + KC - 1) / KC;
    KC = (nsteps + npass - 1) / npass;
  }
  const size_t tsz = dense ? (size_t)L.mtr * L.nt * 128 : 0;            // doubles of fragment scratch per matrix
  const size_t cmsz = (size_t)dr * dr * 2;                               // doubles per prefactor matrix
  const int dp = (d + 1) & ~1;
  const size_t per_traj = (size_t)KC * sizeof(double) * (cmsz + 2 + 8 + 2 * d + tsz + (hconst ? 0 : 4 * hsz + 4 * dp)) +
                          sizeof(double) * (8 * d + 4);                       // + path state of the sGDML stage kernels
  size_t budget = (size_t)6 << 30;
  if (const char *s = getenv("SC_CHUNK_SCRATCH_MB")) budget = (size_t)atol(s) << 20;
  long long ntb = (long long)(budget / per_traj);
  ntb = (ntb / sm) * sm;
  if (ntb < sm) ntb = sm;
  if (ntb > n) ntb = n;
  if (int rc = ensure_chunk_scratch(e, per_traj * (size_t)ntb + 1024, st)) return rc;
  double *base = reinterpret_cast<double *>(e->chunk_scratch);
  double2 *cm = reinterpret_cast<double2 *>(base);            base += (size_t)KC * ntb * cmsz;
  double2 *det = reinterpret_cast<double2 *>(base);           base += (size_t)KC * ntb * 2;
  double *aux = base;                                         base += (size_t)KC * ntb * 8;
  double *qp = base;                                          base += (size_t)KC * ntb * 2 * d;
  double *T = dense ? base : nullptr;                         base += (size_t)KC * ntb * tsz;
  double *hs = hconst ? e->stream_const : base;                 base += hconst ? 0 : (size_t)KC * ntb * 4 * hsz;
  double *hd = base;                                            base += hconst ? 0 : (size_t)KC * ntb * 4 * dp;
  double *pst = base;                                           // sGDML: path state, stage positions, V, grad
  size_t ngroups = 0;
  for (long long t0 = 0; t0 < n; t0 += ntb) ngroups += (size_t)((std::min<long long>(ntb, n - t0) + 127) / 128);
  if (int rc = ensure_partials(e, ngroups * nsteps * 5, st)) return rc;
  StreamArgs A;
  A.hs = hs;
  A.hs_const = hconst ? 1 : 0;
  A.L1p = dense ? e->stream_const + hsz : nullptr;
  A.L2p = dense ? e->stream_const + 2 * hsz : nullptr;
  A.cm = cm;
  A.T = T;
  A.skip_rk4 = 0;
  for (int s0 = 0; s0 < nsteps; s0 += KC) {
    const int ks = std::min(KC, nsteps - s0);
    size_t g0 = 0;
    for (long long t0 = 0; t0 < n; t0 += ntb) {
      const int nt = (int)std::min<long long>(ntb, n - t0);
      timing_mark(e, TS_PATH, st);
      bool aux_done = false;
      if (P.type == POT_HARMONIC) {
        const size_t psm = sizeof(double) * ((size_t)d * (d | 1) + PATH_WARPS * ((d + 1) & ~1));
        CU(cudaFuncSetAttribute(k_path_harmonic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        k_path_harmonic<<<(nt + PATH_WARPS - 1) / PATH_WARPS, 32 * PATH_WARPS, psm, st>>>(e->dev, P, h, ks, (int)t0, nt, qp, aux);
      } else if (P.type == POT_ROTATED_MORSE) {
        const size_t psm = sizeof(double) * ((size_t)d * (d | 1) + PATH_WARPS * 2 * ((d + 1) & ~1));
        CU(cudaFuncSetAttribute(k_path_rotated, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        k_path_rotated<<<(nt + PATH_WARPS - 1) / PATH_WARPS, 32 * PATH_WARPS, psm, st>>>(e->dev, P, h, ks, (int)t0, nt, qp, aux, hd);
      } else if (P.type == POT_GDML) {
        // sGDML: E, grad, Hessian of every stage by k_gdml_eval; the Hessian lands in the stream image of its stage
        double *rbuf = pst + (size_t)6 * d * nt + 2 * nt, *Vb = rbuf + (size_t)d * nt, *gb = Vb + nt;
        const int blk = (nt + 127) / 128;
        k_gstage_begin<<<blk, 128, 0, st>>>(e->dev, (int)t0, nt, pst, rbuf);
        CU(cudaGetLastError());
        for (int step = 0; step < ks; ++step)
          for (int sg = 1; sg <= 4; ++sg) {
            if (launch_gdml_eval(P, nt, rbuf, Vb, gb, hs + ((size_t)(step * 4 + sg - 1) * nt) * hsz, st, L.ldh, hsz))
              return fail(SC_ERR_UNSUPPORTED, "sGDML model outside the kernel's envelope");
            k_gstage_adv<<<blk, 128, 0, st>>>(e->dev, P, h, sg, step, step + 1 == ks ? 1 : 0, (int)t0, nt, pst, Vb, gb, rbuf, qp, aux);
            CU(cudaGetLastError());
            e->launches += 2;
          }
      } else if ((P.type == POT_MORSE || P.type == POT_NONHARMONIC) && e->dev.diag && d <= 64) {
        k_qp_path<<<(nt + 3) / 4, 128, 0, st>>>(e->dev, P, h, ks, (int)t0, nt, hd, aux);      // overlap terms included
        aux_done = true;
      } else if (P.type == POT_MORSE || P.type == POT_NONHARMONIC) {
        k_path_separable<<<(nt + PATH_WARPS - 1) / PATH_WARPS, 32 * PATH_WARPS, 0, st>>>(e->dev, P, h, ks, (int)t0, nt, qp, aux, hd);
      } else {
        return fail(SC_ERR_UNSUPPORTED, "stream pipeline: potential type %d", P.type);
      }
      CU(cudaGetLastError());
      if (!aux_done) {
        const long long items = (long long)ks * nt;
        const int grid = (int)std::min<long long>((items + 7) / 8, (long long)sm * 8);
        k_aux_terms<<<grid, 256, 0, st>>>(e->dev, ks, (int)t0, nt, qp, aux);
        CU(cudaGetLastError());
      }
      if (!hconst && P.type != POT_GDML) {
        timing_mark(e, TS_POT, st);
        CU(launch_expand(P, P.type == POT_ROTATED_MORSE ? 1 : 0, ks, nt, hd, hs, sm, st));
        e->launches += 1;
      }
      timing_mark(e, TS_RK4, st);
      CU(launch_stream((long long)nt * L.ngroups, sm, e->dev, P, h, ks, (int)t0, nt, A, L, st));
      if (dense) {
        timing_mark(e, TS_RMULT, st);
        CU(launch_rmult(e->dev, (long long)ks * nt, T, cm, sm, st));
      }
      timing_mark(e, TS_LU, st);
      CU(launch_lu_batch(cm, dr, ks * nt, det, sm, 0, st));
      timing_mark(e, TS_FINISH, st);
      const int nblk = (nt + 127) / 128;
      k_hk_finish<<<nblk, 128, 0, st>>>(e->dev, (int)t0, nt, ks, s0, nsteps, det, aux, e->partials + g0 * nsteps * 5);
      CU(cudaGetLastError());
      timing_mark(e, TS_END, st);
      g0 += nblk;
      e->launches += dense ? 6 : 5;
    }
  }
  k_reduce_partials<<<nsteps, 160, 0, st>>>(e->partials, (int)ngroups, nsteps, 1.0 / (double)e->ntraj_norm, 1.0 / (double)n, out_dev);
  CU(cudaGetLastError());
  e->launches += 1;
  e->kernel_name = dense ? (dr > 64 ? "k_rk4_stream+k_rmult+k_lu_big+k_hk_finish"
                                    : dr > 24 ? "k_rk4_stream+k_rmult+k_lu_mma+k_hk_finish" : "k_rk4_stream+k_rmult+k_lu_warp+k_hk_finish")
                         : (dr > 64 ? "k_rk4_stream+k_lu_big+k_hk_finish"
                                    : dr > 24 ? "k_rk4_stream+k_lu_mma+k_hk_finish" : "k_rk4_stream+k_lu_warp+k_hk_finish");
  return SC_OK;
}

static int run_hk_kernel(sc_engine *e, const PotDev &P, double h, int nsteps, int mode, double *out_dev, cudaStream_t st,
                         bool allow_mma = true) {
  LaunchPlan pl;
  if (allow_mma && getenv("SC_NO_MMA")) allow_mma = false;  // diagnostics: force the DFMA kernel
  // (the per-step snapshots of the fused Walton-Manolopoulos launches, d <= 16, are written by k_hk_small / k_hk_generic only)
  const bool snapshots = e->dev.snap != nullptr;
  if (allow_mma && !snapshots && mode == MODE_STEP && !getenv("SC_NO_CHUNK") && !e->dense_engine && !getenv("SC_DENSE_ENGINE") &&
      chunk_supported(e->dev, P))
    return run_hk_chunked(e, P, h, nsteps, out_dev, st);
  if (allow_mma && !snapshots && mode == MODE_STEP && !getenv("SC_NO_STREAM") && stream_supported(e->dev, P))
    return run_hk_stream(e, P, h, nsteps, out_dev, st);
  if (e->dev.d > 62) {                       // the set-up / read-out modes of k_hk_generic do not fit in shared memory
    if (mode == MODE_INIT || mode == MODE_TRACK) return run_prefactor_stream(e, mode, st);
    if (mode == MODE_CORR) return run_corr_now(e, out_dev, st);
    return fail(SC_ERR_UNSUPPORTED, "no fused step kernel for this potential at d = %d; use the stage interface", e->dev.d);
  }
  if (allow_mma && mode == MODE_STEP && !getenv("SC_NO_SMALL") && small_supported(e->dev, P)) {
    // register-resident column kernel for the small systems (sc_small.cuh); one partial row per warp and step
    int nwarps = 0;
    CU(launch_small(e->sm_count, e->dev, P, h, nsteps, nullptr, nwarps, true, st));
    const size_t need = (size_t)nwarps * nsteps * 5;
    if (int rc = ensure_partials(e, need, st)) return rc;
    CU(cudaMemsetAsync(e->partials, 0, sizeof(double) * need, st));
    CU(launch_small(e->sm_count, e->dev, P, h, nsteps, e->partials, nwarps, false, st));
    e->kernel_name = "k_hk_small";
    k_reduce_partials<<<nsteps, 160, 0, st>>>(e->partials, nwarps, nsteps, 1.0 / (double)e->ntraj_norm,
                                              1.0 / (double)e->dev.n, out_dev);
    CU(cudaGetLastError());
    e->launches += 2;
    return SC_OK;
  }
  if (int rc = plan_launch(e, pl)) return rc;
  const int nrows = (mode == MODE_STEP) ? nsteps : 1;
  const int ngroups = pl.grid * pl.groups_per_cta;
  const size_t need = (size_t)ngroups * nrows * 5;
  if (need > e->partials_cap) {
    CU(cudaStreamSynchronize(st));
    if (e->partials) cudaFree(e->partials);
    e->partials = nullptr;
    CU(cudaMalloc(&e->partials, sizeof(double) * need));
    e->partials_cap = need;
  }
  if (mode != MODE_INIT && mode != MODE_TRACK) CU(cudaMemsetAsync(e->partials, 0, sizeof(double) * need, st));
  cudaError_t ce = cudaSuccess;
  {
    e->kernel_name = "k_hk_generic";
#define CASE(T, E_) ce = launch_generic<T, E_>(pl, e->dev, P, h, nsteps, mode, e->partials, st)
    if (pl.tpt == 32) {
      if (pl.ept <= 2) CASE(32, 2);
      else if (pl.ept <= 4) CASE(32, 4);
      else if (pl.ept <= 9) CASE(32, 9);
      else CASE(32, 16);
    } else if (pl.tpt == 256) {
      if (pl.ept <= 8) CASE(256, 8);
      else CASE(256, 16);
    } else {
      CASE(320, 26);
    }
#undef CASE
  }
  if (ce != cudaSuccess) return fail(SC_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(ce));
  e->launches += 1;
  if (mode != MODE_INIT && mode != MODE_TRACK) {
    k_reduce_partials<<<nrows, 160, 0, st>>>(e->partials, ngroups, nrows, 1.0 / (double)e->ntraj_norm,
                                             1.0 / (double)e->dev.n, out_dev);
    CU(cudaGetLastError());
    e->launches += 1;
  }
  return SC_OK;
}

static int ensure_corr(sc_engine *e, int nsteps, cudaStream_t st) {
  const size_t need = (size_t)nsteps * 5;
  if (need > e->corr_cap) {
    CU(cudaStreamSynchronize(st));
    if (e->corr_dev) cudaFree(e->corr_dev);
    e->corr_dev = nullptr;
    CU(cudaMalloc(&e->corr_dev, sizeof(double) * need));
    e->corr_cap = need;
  }
  return SC_OK;
}

extern "C" int sc_engine_set_ensemble(sc_engine *e, int n, long long ntraj_norm, const double *zi, const double *probi,
                                      void *stream) {
  if (!e || !zi || !probi) return fail(SC_ERR_INVALID, "null argument");
  if (n < 1) return fail(SC_ERR_INVALID, "need at least one trajectory");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EngDev &D = e->dev;
  const int d = D.d;
  CU(cudaStreamSynchronize(st));
  e->ntraj_norm = ntraj_norm > 0 ? ntraj_norm : n;
  double *zt = const_cast<double *>(D.zt);
  double2 *wvi = const_cast<double2 *>(D.wvi);
  // a new ensemble of the same size (the next repetition of a run) reuses the device buffers: allocating and freeing
  // 116 KB per trajectory costs more than initialising them
  const bool reuse = !e->cfg.wm && e->ens_n == n && !e->ens.ptrs.empty() && !e->stage_buf;
  if (!reuse) {
    e->ens.release();  // the previous ensemble
    e->stage_buf = nullptr;
    e->ens_n = 0;
    D.n = n;
    D.qps = (2 * d + 1 + 1) & ~1;
    D.rs = D.qps + 4 * d * d;
    CU(e->ens.alloc((size_t)n * D.rs, &D.rec));
    CU(e->ens.alloc((size_t)n * 2 * d, &zt));
    CU(e->ens.alloc((size_t)n, &wvi));
    CU(e->ens.alloc((size_t)n, &D.c2));
    CU(e->ens.alloc((size_t)n, &D.c));
    CU(e->ens.alloc((size_t)n, &D.sign));
    D.zt = zt;
    D.wvi = wvi;
    e->ens_n = n;
  }
  CU(cudaMemsetAsync(D.c2, 0, sizeof(double2) * n, st));
  CU(cudaMemsetAsync(D.c, 0, sizeof(double2) * n, st));
  CU(cudaMemsetAsync(D.sign, 0, sizeof(double) * n, st));
  const double inv2pid = std::pow(2.0 * M_PI, -(double)d);
  k_init_records<<<(n < e->sm_count * 16 ? n : e->sm_count * 16), 128, 0, st>>>(D, zi, probi, e->oiA, e->oiB, e->oiC, e->cfg.oi0_fac, inv2pid, zt, wvi);
  CU(cudaGetLastError());
  e->launches += 1;
  // prefactor at t = 0 initialises the branch trackers (propagators.py:628-631)
  PotDev none = PotDev();
  none.d = d;
  none.imass = D.q0;  // never dereferenced beyond d entries in MODE_INIT
  // at t = 0 every trajectory has Mqq = Mpp = 1, Mqp = Mpq = 0: the prefactor matrix (propagators.py:969-994) and its
  // determinant are the same for the whole ensemble -> one LU, then broadcast (bit-identical to n separate LUs)
  {
    const int n_all = D.n;
    D.n = 1;
    const int rc = run_hk_kernel(e, none, 0.0, 0, MODE_INIT, nullptr, st);
    D.n = n_all;
    if (rc) return rc;
    k_broadcast_prefactor<<<(n_all + 255) / 256, 256, 0, st>>>(D.c2, D.c, D.sign, n_all);
    CU(cudaGetLastError());
    e->launches += 1;
  }
  if (e->cfg.wm) {
    if (int rc = wm_alloc(e->wm, e->ens, D, D.q0, D.p0, probi, e->sm_count, st)) return rc;
    if (int rc = wm_launch(e->wm, e->dev, WM_INIT, 0.0, nullptr, nullptr, st)) return rc;
    e->launches += 2;
  }
  return SC_OK;
}

// initial_conditions: importance sampling of (qi, pi) ~ |<qi,pi,Gi|q0,p0,G0>|^2 on the device (propagators.py:533-555);
// iLq, iLp (d' x d) are the blocks of Lz^-1 computed on the host exactly as the reference does (:506-515), detLz as :531.
// The ensemble is a pure function of (seed, index0 + local index): ranks pass their shard's first global index.
extern "C" int sc_engine_sample_ensemble(sc_engine *e, int n, long long index0, unsigned long long seed, const double *iLq_host,
                                         const double *iLp_host, double detLz, double *zi_dev, double *probi_dev, void *stream) {
  if (!e || n < 1 || !iLq_host || !iLp_host || !zi_dev || !probi_dev) return fail(SC_ERR_INVALID, "sample_ensemble(): bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d = e->dev.d, dr = e->dev.dr;
  const size_t m = (size_t)dr * d;
  if (2 * m > e->stage_in_cap) {
    CU(cudaStreamSynchronize(st));
    if (e->stage_in) cudaFree(e->stage_in);
    e->stage_in = nullptr;
    e->stage_in_cap = 0;
    CU(cudaMalloc(&e->stage_in, sizeof(double) * 2 * m));
    e->stage_in_cap = 2 * m;
  }
  CU(cudaMemcpyAsync(e->stage_in, iLq_host, sizeof(double) * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(e->stage_in + m, iLp_host, sizeof(double) * m, cudaMemcpyHostToDevice, st));
  const double pfac = detLz * std::pow(2.0 * M_PI, -(double)d);
  k_sample_ensemble<<<(n + 3) / 4, 128, 0, st>>>(d, dr, n, index0, seed, e->stage_in, e->stage_in + m, e->dev.q0, e->dev.p0, pfac,
                                                zi_dev, probi_dev);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));            // the host arrays may go away; the staging buffer may be reused
  e->launches += 1;
  return SC_OK;
}

extern "C" int sc_engine_set_ensemble_host(sc_engine *e, int n, long long ntraj_norm, const double *zi_host,
                                           const double *probi_host, void *stream) {
  if (!e || !zi_host || !probi_host) return fail(SC_ERR_INVALID, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nz = (size_t)2 * e->dev.d * n;
  if (nz + n > e->stage_in_cap) {                        // staging buffer kept for the next repetition
    CU(cudaStreamSynchronize(st));
    if (e->stage_in) cudaFree(e->stage_in);
    e->stage_in = nullptr;
    e->stage_in_cap = 0;
    CU(cudaMalloc(&e->stage_in, sizeof(double) * (nz + n)));
    e->stage_in_cap = nz + n;
  }
  double *zi = e->stage_in, *pr = zi + nz;
  CU(cudaMemcpyAsync(zi, zi_host, sizeof(double) * nz, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(pr, probi_host, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  return sc_engine_set_ensemble(e, n, ntraj_norm, zi, pr, stream);
}

static int check_step_args(sc_engine *e, const sc_potential *pot) {
  if (!e || !pot) return fail(SC_ERR_INVALID, "null argument");
  if (e->dev.n < 1) return fail(SC_ERR_INVALID, "initial_conditions / sc_engine_set_ensemble has not been called");
  if (pot->dev.d != e->dev.d) return fail(SC_ERR_INVALID, "potential has wrong dimensions");
  return SC_OK;
}

extern "C" int sc_engine_step_dev(sc_engine *e, const sc_potential *pot, double dt, int nsteps, double *corr_dev,
                                  void *stream) {
  if (int rc = check_step_args(e, pot)) return rc;
  if (nsteps < 1) return SC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = set_nac(e, pot->n1.data(), st)) return rc;
  if (pot->dev.type == POT_GDML && !stream_supported(e->dev, pot->dev))
    return fail(SC_ERR_UNSUPPORTED, "sGDML potentials with d < 17 or d > 64 run through the stage interface");
  if (!corr_dev) {
    if (int rc = ensure_corr(e, nsteps, st)) return rc;
    corr_dev = e->corr_dev;
  }
  if (!e->cfg.wm) return run_hk_kernel(e, pot->dev, dt, nsteps, MODE_STEP, corr_dev, st);
  if (e->dev.d <= 16 && !getenv("SC_WM_UNFUSED")) {
    // Walton-Manolopoulos, K steps per launch: k_hk_generic advances the trajectories KC steps and snapshots every new time
    // (record, sqrt(det), sign); ONE k_wm_fused launch evaluates the Filinov-smoothed prefactor pieces and the contributions
    // of all KC times (branch trackers walked in time order per trajectory)
    WMState &w = e->wm;
    const int n = e->dev.n;
    const size_t per_step = sizeof(double) * ((size_t)n * e->dev.rs + 3 * (size_t)n);
    int KC = (int)std::max<size_t>(1, std::min<size_t>((size_t)nsteps, ((size_t)4 << 30) / per_step));
    if (const char *s = getenv("SC_CHUNK_K")) KC = std::max(1, std::min(KC, atoi(s)));
    const size_t ngr = (size_t)w.grid * w.groups;
    const size_t need = per_step * KC + sizeof(double) * (ngr * KC * 4 + (size_t)KC * 5) + 256;
    if (need > w.snap_cap) {
      CU(cudaStreamSynchronize(st));
      if (w.snap) cudaFree(w.snap);
      w.snap = nullptr;
      w.snap_cap = 0;
      CU(cudaMalloc(&w.snap, need));
      w.snap_cap = need;
    }
    double *sb = reinterpret_cast<double *>(w.snap);
    double *snap = sb;                                              sb += (size_t)KC * n * e->dev.rs;
    double2 *snap_c = reinterpret_cast<double2 *>(sb);              sb += (size_t)KC * n * 2;
    double *snap_sign = sb;                                         sb += (size_t)KC * n;
    double *wpart = sb;                                             sb += ngr * KC * 4;
    double *hkrows = sb;
    const int threads = w.tpt * w.groups;
    for (int s0 = 0; s0 < nsteps; s0 += KC) {
      const int ks = std::min(KC, nsteps - s0);
      e->dev.snap = snap;
      e->dev.snap_c = snap_c;
      e->dev.snap_sign = snap_sign;
      int rc = run_hk_kernel(e, pot->dev, dt, ks, MODE_STEP, hkrows, st);
      EngDev Dsnap = e->dev;
      e->dev.snap = nullptr;
      e->dev.snap_c = nullptr;
      e->dev.snap_sign = nullptr;
      if (rc) return rc;
      CU(cudaMemsetAsync(wpart, 0, sizeof(double) * ngr * ks * 4, st));
      if (w.tpt == 32 && e->dev.d == 5 && e->dev.dr == 5 && !getenv("SC_WM_RUNTIME_D")) {
        // BASELINE configs[1]: 5 modes, full-rank widths -- dimension folded at compile time
        CU(cudaFuncSetAttribute(k_wm_fused<32, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
        k_wm_fused<32, 5><<<w.grid, threads, w.smem, st>>>(Dsnap, w.dev, w.L, ks, wpart);
      } else if (w.tpt == 32) {
        CU(cudaFuncSetAttribute(k_wm_fused<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
        k_wm_fused<32><<<w.grid, threads, w.smem, st>>>(Dsnap, w.dev, w.L, ks, wpart);
      } else {
        CU(cudaFuncSetAttribute(k_wm_fused<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
        k_wm_fused<128><<<w.grid, threads, w.smem, st>>>(Dsnap, w.dev, w.L, ks, wpart);
      }
      CU(cudaGetLastError());
      k_wm_reduce_k<<<ks, 160, 0, st>>>(wpart, (int)ngr, ks, 1.0 / (double)e->ntraj_norm, hkrows, corr_dev + 5 * s0);
      CU(cudaGetLastError());
      e->launches += 3;
    }
    e->kernel_name = !strcmp(e->kernel_name, "k_hk_small") ? "k_hk_small+k_wm_fused" : "k_hk_generic+k_wm_fused";
    return SC_OK;
  }
  // larger d: the HK pipeline advances the trajectories one step at a time, the WM kernel evaluates the Filinov-smoothed
  // prefactor pieces and the WM contributions of every new time
  for (int k = 0; k < nsteps; ++k) {
    if (int rc = run_hk_kernel(e, pot->dev, dt, 1, MODE_STEP, e->wm.scratch5, st)) return rc;
    if (int rc = wm_launch(e->wm, e->dev, WM_STEP, 1.0 / (double)e->ntraj_norm, corr_dev + 5 * k, e->wm.scratch5, st)) return rc;
    e->launches += 2;
  }
  e->kernel_name = e->wm.global_ws ? "HK step pipeline+k_wm_global" : "HK step pipeline+k_wm";
  return SC_OK;
}

extern "C" int sc_engine_step(sc_engine *e, const sc_potential *pot, double dt, int nsteps, double *corr_host,
                              void *stream) {
  if (int rc = check_step_args(e, pot)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = ensure_corr(e, nsteps, st)) return rc;
  if (int rc = sc_engine_step_dev(e, pot, dt, nsteps, e->corr_dev, stream)) return rc;
  if (corr_host) {
    CU(cudaMemcpyAsync(corr_host, e->corr_dev, sizeof(double) * 5 * nsteps, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return SC_OK;
}

static int correlations_impl(sc_engine *e, const PotDev &P, const double *n1, double *out_host, cudaStream_t st) {
  if (int rc = set_nac(e, n1, st)) return rc;
  if (int rc = ensure_corr(e, 1, st)) return rc;
  if (!e->cfg.wm) {
    if (int rc = run_hk_kernel(e, P, 0.0, 1, MODE_CORR, e->corr_dev, st)) return rc;
  } else {
    if (int rc = wm_launch(e->wm, e->dev, WM_CORR, 1.0 / (double)e->ntraj_norm, e->corr_dev, nullptr, st)) return rc;
    e->launches += 2;
  }
  CU(cudaMemcpyAsync(out_host, e->corr_dev, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SC_OK;
}

extern "C" int sc_engine_correlations(sc_engine *e, const sc_potential *pot, double *out_host, void *stream) {
  if (int rc = check_step_args(e, pot)) return rc;
  if (!out_host) return fail(SC_ERR_INVALID, "null argument");
  return correlations_impl(e, pot->dev, pot->n1.data(), out_host, static_cast<cudaStream_t>(stream));
}

extern "C" int sc_engine_correlations_n1(sc_engine *e, const double *n1_host, double *out_host, void *stream) {
  if (!e || !n1_host || !out_host) return fail(SC_ERR_INVALID, "null argument");
  if (e->dev.n < 1) return fail(SC_ERR_INVALID, "no ensemble");
  PotDev none = PotDev();
  none.d = e->dev.d;
  none.imass = e->dev.q0;
  none.n1 = nullptr;
  // the WM kernel reads n1 from device memory: stage it through the engine's cache
  return correlations_impl(e, none, n1_host, out_host, static_cast<cudaStream_t>(stream));
}

// autocorrelation / ic_correlation at the current time with POSITION-DEPENDENT couplings tau1(r), tau2(r)
// (propagators.py:868-909 in full generality; Herman-Kluk).  n1Q_dev, n1q_dev: (d, n) = -hbar^2 tau1 / m at the current and at
// the initial positions; n2Q_dev, n2q_dev: (n) = -hbar^2 / 2 sum_k tau2_k / m_k.  out_host: 4 doubles (no phase factor).
extern "C" int sc_engine_correlations_general(sc_engine *e, const double *n1Q_dev, const double *n1q_dev, const double *n2Q_dev,
                                              const double *n2q_dev, double *out_host, void *stream) {
  if (!e || e->dev.n < 1 || !n1Q_dev || !n1q_dev || !n2Q_dev || !n2q_dev || !out_host) return fail(SC_ERR_INVALID, "null argument / no ensemble");
  if (e->cfg.wm) return fail(SC_ERR_UNSUPPORTED, "position-dependent couplings: Herman-Kluk propagator only");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = e->dev.n;
  int grid = std::min((n + 7) / 8, e->sm_count * 8);
  if (grid < 1) grid = 1;
  if (int rc = ensure_partials(e, (size_t)grid * 5, st)) return rc;
  if (int rc = ensure_corr(e, 1, st)) return rc;
  k_corr_general<<<grid, 256, 0, st>>>(e->dev, e->d_R, e->d_G0iG, n1Q_dev, n1q_dev, n2Q_dev, n2q_dev, e->partials);
  CU(cudaGetLastError());
  k_reduce_partials<<<1, 160, 0, st>>>(e->partials, grid, 1, 1.0 / (double)e->ntraj_norm, 1.0 / (double)n, e->corr_dev);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out_host, e->corr_dev, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  e->launches += 2;
  return SC_OK;
}

extern "C" int sc_engine_get_state(sc_engine *e, double *y, void *stream) {
  if (!e || !y || e->dev.n < 1) return fail(SC_ERR_INVALID, "no ensemble / null argument");
  k_export_state<<<e->sm_count * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(e->dev, y, 1);
  CU(cudaGetLastError());
  e->launches += 1;
  return SC_OK;
}

extern "C" int sc_engine_set_state(sc_engine *e, const double *y, void *stream) {
  if (!e || !y || e->dev.n < 1) return fail(SC_ERR_INVALID, "no ensemble / null argument");
  k_export_state<<<e->sm_count * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(e->dev, const_cast<double *>(y), 0);
  CU(cudaGetLastError());
  e->launches += 1;
  return SC_OK;
}

__global__ void k_fill(double *x, size_t n, double v) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

extern "C" int sc_engine_get_prefactor(sc_engine *e, double *c, double *c2, double *signs, void *stream) {
  if (!e || e->dev.n < 1) return fail(SC_ERR_INVALID, "no ensemble");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = e->dev.n;
  if (c) CU(cudaMemcpyAsync(c, e->dev.c, sizeof(double2) * n, cudaMemcpyDeviceToDevice, st));
  if (c2) CU(cudaMemcpyAsync(c2, e->dev.c2, sizeof(double2) * n, cudaMemcpyDeviceToDevice, st));
  if (signs) {
    CU(cudaMemcpyAsync(signs, e->dev.sign, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    if (e->cfg.wm) {
      CU(cudaMemcpyAsync(signs + n, e->wm.dev.signA, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
      CU(cudaMemcpyAsync(signs + 2 * n, e->wm.dev.signM, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    } else {
      k_fill<<<(2 * n + 255) / 256, 256, 0, st>>>(signs + n, 2 * (size_t)n, 1.0);   // no "detA" / "detM" trackers: +1
      CU(cudaGetLastError());
    }
  }
  return SC_OK;
}

// ------------------------------------------------------------------ wavefunction diagnostics (sc_gauss.cuh) ---------
namespace {
cudaError_t launch_gauss_sum(int n_bra, int n_ket, int kp, const double *a, const double *alpha, const double *gamma,
                             const double *r, const double *s, const double *alphaJ, const double *beta, const double2 *coef,
                             double2 *out, cudaStream_t st) {
  const size_t smem = gs_smem_bytes(kp);
  cudaError_t ce = cudaFuncSetAttribute(k_gauss_sum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce != cudaSuccess) return ce;
  k_gauss_sum<<<(n_bra + GS_TI - 1) / GS_TI, GS_THREADS, smem, st>>>(n_bra, n_ket, kp, a, alpha, gamma, r, s, alphaJ, beta, coef, out);
  return cudaGetLastError();
}
}  // namespace

// Walton-Manolopoulos diagnostics (propagators.py:1391-1575): one WM_DIAG pass of k_wm rebuilds CQQ, CqQ, PI_Q, det A of the
// current state and leaves the coefficients v_n (eqn 75) and the per-trajectory matrices of the wavefunction / norm kernels in
// trajectory-minor arrays
static int wm_diag_pass(sc_engine *e, cudaStream_t st, double **partials_out) {
  WMState &w = e->wm;
  const int n = e->dev.n, d = e->dev.d, dr = e->dev.dr;
  if (d > WMD_MAX) return fail(SC_ERR_UNSUPPORTED, "Walton-Manolopoulos diagnostics: d = %d > %d", d, WMD_MAX);
  const size_t per = 2 * (1 + (size_t)d * d + (size_t)dr * d + (size_t)dr * dr + d + dr) + d;      // doubles per trajectory
  const size_t need = per * n + 2 * (size_t)e->sm_count * 8 + 16;
  if (need > w.diag_cap) {
    CU(cudaStreamSynchronize(st));
    if (w.diag) cudaFree(w.diag);
    w.diag = nullptr;
    w.diag_cap = 0;
    CU(cudaMalloc(&w.diag, sizeof(double) * need));
    w.diag_cap = need;
  }
  double2 *b2 = reinterpret_cast<double2 *>(w.diag);
  w.dev.dg_v = b2;                  b2 += n;
  w.dev.dg_CQQ = b2;                b2 += (size_t)d * d * n;
  w.dev.dg_UC = b2;                 b2 += (size_t)dr * d * n;
  w.dev.dg_CP = b2;                 b2 += (size_t)dr * dr * n;
  w.dev.dg_D = b2;                  b2 += (size_t)d * n;
  w.dev.dg_DP = b2;                 b2 += (size_t)dr * n;
  w.dev.dg_Q = reinterpret_cast<double *>(b2);
  if (partials_out) *partials_out = w.dev.dg_Q + (size_t)d * n;
  w.dev.diag_inv_norm = 1.0 / (double)e->ntraj_norm;
  const int threads = w.tpt * w.groups;
  if (w.tpt == 32) {
    CU(cudaFuncSetAttribute(k_wm<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
    k_wm<32><<<w.grid, threads, w.smem, st>>>(e->dev, w.dev, w.L, WM_DIAG, w.partials);
  } else {
    CU(cudaFuncSetAttribute(k_wm<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.smem));
    k_wm<128><<<w.grid, threads, w.smem, st>>>(e->dev, w.dev, w.L, WM_DIAG, w.partials);
  }
  CU(cudaGetLastError());
  e->launches += 1;
  return SC_OK;
}

// expansion coefficients v_i of the frozen-Gaussian wavefunction (HermanKlukPropagator.coefficients, propagators.py:657-686;
// WaltonManolopoulosPropagator.coefficients, propagators.py:1391-1432)
extern "C" int sc_engine_coefficients(sc_engine *e, double *v_dev, void *stream) {
  if (!e || e->dev.n < 1 || !v_dev) return fail(SC_ERR_INVALID, "no ensemble");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (e->cfg.wm) {
    if (int rc = wm_diag_pass(e, st, nullptr)) return rc;
    CU(cudaMemcpyAsync(v_dev, e->wm.dev.dg_v, sizeof(double2) * e->dev.n, cudaMemcpyDeviceToDevice, st));
    return SC_OK;
  }
  const int n = e->dev.n;
  k_coefficients<<<(n + 255) / 256, 256, 0, st>>>(e->dev, 1.0 / (double)e->ntraj_norm, reinterpret_cast<double2 *>(v_dev));
  CU(cudaGetLastError());
  return SC_OK;
}

// WaltonManolopoulosPropagator.norm (propagators.py:1484-1575): all pairs, a (d' x d') complex inverse + determinant per pair
static int wm_norm(sc_engine *e, double *norm2_host, cudaStream_t st) {
  double *partials = nullptr;
  if (int rc = wm_diag_pass(e, st, &partials)) return rc;
  const int n = e->dev.n;
  int grid = std::min(n, e->sm_count * 8);
  k_wm_norm<<<grid, 128, 0, st>>>(e->dev.d, e->dev.dr, n, e->wm.dev, partials);
  CU(cudaGetLastError());
  double *res = partials + 2 * (size_t)grid;
  k_wm_norm_reduce<<<1, 64, 0, st>>>(partials, grid, res);
  CU(cudaGetLastError());
  double h[2];
  CU(cudaMemcpyAsync(h, res, sizeof(h), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  norm2_host[0] = h[0];
  norm2_host[1] = h[1];
  e->launches += 2;
  return SC_OK;
}

// |psi|^2 = sum_ij conj(v_i) <q_i,p_i,Gt|q_j,p_j,Gt> v_j (HermanKlukPropagator.norm, propagators.py:734-782).  A, B, C: the
// (d x d) matrices Gt (2 Gt)^+ Gt, (2 Gt)^+, Gt (2 Gt)^+ of CoherentStatesOverlap(Gt, Gt) (propagators.py:174-179), fac its
// normalisation factor (:230).
// Sharded ensembles: every rank PACKS the ket vectors of its shard (sc_engine_norm_pack), the packs are all-gathered by the
// caller (torch.distributed / NCCL), and every rank sums its (n_local x n_r) BLOCK against each rank's pack
// (sc_engine_norm_block); the caller adds the blocks and all-reduces the two doubles.  Pack layout for n_pad rows:
// [r (n_pad x kp) | s (n_pad x kp) | alphaJ (n_pad) | beta (n_pad) | coef (n_pad x 2)],  kp = (2 d + 3) & ~3.
// The bra-side vectors stay in a persistent scratch of the engine (no allocation per call after the first).
static int norm_scratch(sc_engine *e, size_t doubles, cudaStream_t st) {
  if (doubles > e->diag_cap) {
    CU(cudaStreamSynchronize(st));
    if (e->diag_scratch) cudaFree(e->diag_scratch);
    e->diag_scratch = nullptr;
    e->diag_cap = 0;
    CU(cudaMalloc(&e->diag_scratch, sizeof(double) * doubles));
    e->diag_cap = doubles;
  }
  return SC_OK;
}

extern "C" int sc_engine_norm_pack_size(const sc_engine *e, int n_pad, long long *doubles_out) {
  if (!e || !doubles_out || n_pad < 1) return fail(SC_ERR_INVALID, "norm_pack_size(): bad arguments");
  const int kp = (2 * e->dev.d + 3) & ~3;
  *doubles_out = (long long)n_pad * (2 * kp + 4);
  return SC_OK;
}

extern "C" int sc_engine_norm_pack(sc_engine *e, const double *A_host, const double *B_host, const double *C_host, int n_pad,
                                   double *pack_dev, void *stream) {
  if (!e || e->dev.n < 1 || !A_host || !B_host || !C_host || !pack_dev || n_pad < e->dev.n)
    return fail(SC_ERR_INVALID, "norm_pack(): bad arguments");
  if (e->cfg.wm) return fail(SC_ERR_UNSUPPORTED, "sharded norm(): Herman-Kluk propagator only");
  if (gs_smem_bytes((2 * e->dev.d + 3) & ~3) > 227 * 1024) return fail(SC_ERR_UNSUPPORTED, "norm(): d = %d exceeds the all-pairs kernel's tiles (d <= 64)", e->dev.d);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = e->dev.n, d = e->dev.d, kp = (2 * d + 3) & ~3;
  // bra side: mats (3 d^2) | a (n kp) | alpha (n) | gamma (n) | o (2 n) | res (2)
  const size_t m3 = (3 * (size_t)d * d + 1) & ~(size_t)1;
  if (int rc = norm_scratch(e, m3 + (size_t)n * (kp + 4) + 8, st)) return rc;
  double *mats = e->diag_scratch, *a = mats + m3, *alpha = a + (size_t)n * kp, *gamma = alpha + n;
  double *r = pack_dev, *s = r + (size_t)n_pad * kp, *alphaJ = s + (size_t)n_pad * kp, *beta = alphaJ + n_pad;
  double2 *v = reinterpret_cast<double2 *>(beta + n_pad);
  CU(cudaMemcpyAsync(mats, A_host, sizeof(double) * d * d, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(mats + d * d, B_host, sizeof(double) * d * d, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(mats + 2 * d * d, C_host, sizeof(double) * d * d, cudaMemcpyHostToDevice, st));
  k_coefficients<<<(n + 255) / 256, 256, 0, st>>>(e->dev, 1.0 / (double)e->ntraj_norm, v);
  CU(cudaGetLastError());
  k_gauss_prep_norm<<<n, 128, 0, st>>>(e->dev, mats, mats + d * d, mats + 2 * d * d, kp, a, r, s, alpha, beta, gamma);
  CU(cudaGetLastError());
  // the kets' alpha is the bras' alpha (same trajectories)
  CU(cudaMemcpyAsync(alphaJ, alpha, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
  e->launches += 2;
  return SC_OK;
}

// norm2_host[0:2] += fac * sum_{i local} conj(v_i) sum_{j < n_ket} O_ij v_j  for the kets of one pack.  bra_coef_dev: the
// coefficient section of THIS rank's own pack (v_i)
extern "C" int sc_engine_norm_block(sc_engine *e, int n_ket, int n_pad, const double *pack_dev, const double *bra_coef_dev,
                                    double fac, double *norm2_host, void *stream) {
  if (!e || e->dev.n < 1 || !pack_dev || !bra_coef_dev || !norm2_host || n_ket < 0 || n_ket > n_pad)
    return fail(SC_ERR_INVALID, "norm_block(): bad arguments");
  if (n_ket == 0) return SC_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = e->dev.n, d = e->dev.d, kp = (2 * d + 3) & ~3;
  const size_t m3 = (3 * (size_t)d * d + 1) & ~(size_t)1;
  if (e->diag_cap < m3 + (size_t)n * (kp + 4) + 8) return fail(SC_ERR_INVALID, "norm_block(): call sc_engine_norm_pack first");
  double *mats = e->diag_scratch, *a = mats + m3, *alpha = a + (size_t)n * kp, *gamma = alpha + n;
  double2 *o = reinterpret_cast<double2 *>(gamma + n);
  double *res = reinterpret_cast<double *>(o + n);
  const double *r = pack_dev, *s = r + (size_t)n_pad * kp, *alphaJ = s + (size_t)n_pad * kp, *beta = alphaJ + n_pad;
  const double2 *v = reinterpret_cast<const double2 *>(beta + n_pad);
  CU(launch_gauss_sum(n, n_ket, kp, a, alpha, gamma, r, s, alphaJ, beta, v, o, st));
  k_gauss_dot<<<1, 256, 0, st>>>(n, reinterpret_cast<const double2 *>(bra_coef_dev), o, res);
  CU(cudaGetLastError());
  double h[2];
  CU(cudaMemcpyAsync(h, res, sizeof(h), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  norm2_host[0] += fac * h[0];
  norm2_host[1] += fac * h[1];
  e->launches += 2;
  return SC_OK;
}

// single-rank convenience: pack + one block
extern "C" int sc_engine_norm(sc_engine *e, const double *A_host, const double *B_host, const double *C_host, double fac,
                              double *norm2_host, void *stream) {
  if (!e || e->dev.n < 1 || !norm2_host) return fail(SC_ERR_INVALID, "norm(): bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (e->cfg.wm) return wm_norm(e, norm2_host, st);           // A, B, C, fac are Herman-Kluk quantities: unused
  const int n = e->dev.n, kp = (2 * e->dev.d + 3) & ~3;
  const size_t packsz = (size_t)n * (2 * kp + 4);
  if (packsz > e->pack_cap) {
    CU(cudaStreamSynchronize(st));
    if (e->pack_scratch) cudaFree(e->pack_scratch);
    e->pack_scratch = nullptr;
    e->pack_cap = 0;
    CU(cudaMalloc(&e->pack_scratch, sizeof(double) * packsz));
    e->pack_cap = packsz;
  }
  if (int rc = sc_engine_norm_pack(e, A_host, B_host, C_host, n, e->pack_scratch, stream)) return rc;
  norm2_host[0] = norm2_host[1] = 0.0;
  const double *coef = e->pack_scratch + (size_t)n * (2 * kp + 2);
  return sc_engine_norm_block(e, n, n, e->pack_scratch, coef, fac, norm2_host, stream);
}

// psi(x_k) = sum_i v_i <x_k|q_i,p_i,Gt> on nx grid points (HermanKlukPropagator.wavefunction, propagators.py:688-732);
// x_dev: (d, nx) like the reference's argument; fac = (det Gt / pi^rank)^(1/4) (propagators.py:279); phi_dev: c128 (nx)
extern "C" int sc_engine_wavefunction(sc_engine *e, const double *Gt_host, double fac, int nx, const double *x_dev,
                                      double *phi_dev, void *stream) {
  if (!e || e->dev.n < 1 || !Gt_host || !x_dev || !phi_dev || nx < 1) return fail(SC_ERR_INVALID, "wavefunction(): bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (e->cfg.wm) {
    // eqn (75), propagators.py:1434-1482 (Gt_host, fac are Herman-Kluk quantities: unused)
    if (int rc = wm_diag_pass(e, st, nullptr)) return rc;
    k_wm_wavefunction<<<std::min(nx, e->sm_count * 8), 256, 0, st>>>(e->dev.d, e->dev.n, nx, x_dev, e->wm.dev,
                                                                     reinterpret_cast<double2 *>(phi_dev));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    e->launches += 1;
    return SC_OK;
  }
  const int n = e->dev.n, d = e->dev.d, kp = (d + 3) & ~3;
  if (gs_smem_bytes(kp) > 227 * 1024) return fail(SC_ERR_UNSUPPORTED, "wavefunction(): d = %d exceeds the kernel's tiles", d);
  // persistent scratch (pack_scratch): G | a (nx kp) | alpha, gamma (nx) | r, s (n kp) | alphaJ, beta (n) | v (2 n)
  const size_t g2 = ((size_t)d * d + 1) & ~(size_t)1;
  const size_t need = g2 + (size_t)nx * (kp + 2) + (size_t)n * (2 * kp + 4) + 8;
  if (need > e->pack_cap) {
    CU(cudaStreamSynchronize(st));
    if (e->pack_scratch) cudaFree(e->pack_scratch);
    e->pack_scratch = nullptr;
    e->pack_cap = 0;
    CU(cudaMalloc(&e->pack_scratch, sizeof(double) * need));
    e->pack_cap = need;
  }
  double *G = e->pack_scratch, *a = G + g2, *alpha = a + (size_t)nx * kp, *gamma = alpha + nx;
  double *r = gamma + nx, *s = r + (size_t)n * kp, *alphaJ = s + (size_t)n * kp, *beta = alphaJ + n;
  double2 *v = reinterpret_cast<double2 *>(beta + n);
  CU(cudaMemcpyAsync(G, Gt_host, sizeof(double) * d * d, cudaMemcpyHostToDevice, st));
  k_coefficients<<<(n + 255) / 256, 256, 0, st>>>(e->dev, fac / (double)e->ntraj_norm, v);
  CU(cudaGetLastError());
  k_gauss_prep_wf_kets<<<n, 128, 0, st>>>(e->dev, G, kp, r, s, alphaJ, beta);
  CU(cudaGetLastError());
  k_gauss_prep_wf_bras<<<nx, 128, 0, st>>>(d, nx, x_dev, G, kp, a, alpha, gamma);
  CU(cudaGetLastError());
  CU(launch_gauss_sum(nx, n, kp, a, alpha, gamma, r, s, alphaJ, beta, v, reinterpret_cast<double2 *>(phi_dev), st));
  CU(cudaStreamSynchronize(st));
  e->launches += 4;
  return SC_OK;
}

// ------------------------------------------------------------------ FP64 roofline denominator, measured live
// register-resident chains of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) and of DFMA on every SM; best of `reps` runs with
// CUDA events.  out_host[0] = DMMA TFLOP/s, out_host[1] = DFMA TFLOP/s (same kernels as tools/fp64_peak.cu)
namespace {
template <int ILP>
__global__ void k_peak_dmma(double *out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_peak_dfma(double *out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int sc_measure_fp64_peak(double *out_host, int reps, void *stream) {
  if (!out_host) return fail(SC_ERR_INVALID, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int threads = 512, blocks = sms * 2, iters = 20000;
  double *out = nullptr;
  CU(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double best[2] = {0.0, 0.0};
  if (reps < 1) reps = 1;
  for (int which = 0; which < 2; ++which)
    for (int r = 0; r < reps + 1; ++r) {          // first run: warm-up
      CU(cudaEventRecord(e0, st));
      if (which == 0) k_peak_dmma<8><<<blocks, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
      else k_peak_dfma<8><<<blocks, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
      CU(cudaEventRecord(e1, st));
      CU(cudaEventSynchronize(e1));
      float ms = 0.0f;
      CU(cudaEventElapsedTime(&ms, e0, e1));
      const double flop = (which == 0 ? 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32)
                                      : 2.0 * 8 * iters * (double)blocks * threads);
      if (r > 0) best[which] = std::max(best[which], flop / (ms * 1e-3) / 1e12);
    }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  out_host[0] = best[0];
  out_host[1] = best[1];
  return SC_OK;
}

// run-time options of an engine (diagnostics / measurement; defaults are the production dispatch)
extern "C" int sc_engine_set_option(sc_engine *e, const char *name, int value) {
  if (!e || !name) return fail(SC_ERR_INVALID, "null argument");
  if (std::strcmp(name, "dense_engine") == 0) { e->dense_engine = value != 0; return SC_OK; }
  return fail(SC_ERR_INVALID, "unknown option '%s'", name);
}

// per-kernel timing of the chunked path: enable, run sc_engine_step*, then read the accumulated milliseconds
extern "C" int sc_engine_set_timing(sc_engine *e, int on) {
  if (!e) return fail(SC_ERR_INVALID, "null engine");
  e->timing = on != 0;
  e->tev_used = 0;
  for (double &x : e->tms) x = 0.0;
  return SC_OK;
}

static int timing_collect(sc_engine *e) {
  if (e->tev_used) {
    CU(cudaEventSynchronize(e->tev[e->tev_used - 1]));
    CU(cudaDeviceSynchronize());
    for (size_t i = 0; i + 1 < e->tev_used; ++i) {
      const int slot = e->tev_slot[i];
      if (slot < 0 || slot >= SC_TIMING_SLOTS) continue;
      float ms = 0.0f;
      CU(cudaEventElapsedTime(&ms, e->tev[i], e->tev[i + 1]));
      e->tms[slot] += ms;
    }
    e->tev_used = 0;
  }
  return SC_OK;
}

extern "C" int sc_engine_get_timing(sc_engine *e, double *ms4) {
  if (!e || !ms4) return fail(SC_ERR_INVALID, "null argument");
  if (int rc = timing_collect(e)) return rc;
  for (int k = 0; k < 4; ++k) ms4[k] = e->tms[k];
  return SC_OK;
}

extern "C" int sc_engine_get_timing_slots(sc_engine *e, double *ms, int nslots) {
  if (!e || !ms) return fail(SC_ERR_INVALID, "null argument");
  if (int rc = timing_collect(e)) return rc;
  for (int k = 0; k < nslots; ++k) ms[k] = k < SC_TIMING_SLOTS ? e->tms[k] : 0.0;
  return SC_OK;
}

// ------------------------------------------------------------------ generic-potential stage path
#include "sc_stage.cuh"
