/*
 * semiclassical_b200.h -- C ABI of the B200-native semiclassical propagation engine.
 *
 * Drop-in boundary for the hot path of humeniuka/semiclassical's `semi dynamics` task.  The reference has
 * no FFI layer (it is pure Python/torch); the "interface each entry point replaces" is therefore the
 * duck-typed Python protocol of semiclassical/propagators.py and semiclassical/potentials.py.  Every entry
 * point cites the reference method(s) whose numerics it replaces.  The Python host mirror
 * (semiclassical_b200/propagators.py, potentials.py) binds these symbols with ctypes; INTEGRATION.md
 * shows the same stub applied to the reference package itself.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every function returns an int status (SC_OK == 0); sc_last_error() gives the message
 *   - "_dev" pointers are CUDA device pointers (e.g. tensor.data_ptr()), "_host" pointers are host memory
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - real arrays are fp64, complex arrays are interleaved (re, im) fp64 pairs ("c128")
 *   - ensemble arrays use the reference's batch-last layout: zi is (2d, n), y is (2d+4d^2+1, n),
 *     trajectory index contiguous (propagators.py:329-334, 581-603)
 *   - matrices are row-major
 */
#ifndef SEMICLASSICAL_B200_H
#define SEMICLASSICAL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SC_OK 0
#define SC_ERR_INVALID 1      /* bad argument / wrong call order  -> AssertionError in the Python mirror */
#define SC_ERR_CUDA 2         /* CUDA runtime failure             -> RuntimeError */
#define SC_ERR_UNSUPPORTED 3  /* configuration outside the kernels' envelope -> NotImplementedError */

#define SC_ABI_VERSION 1
#define SC_MAX_DIM 96         /* largest number of degrees of freedom: separable / rotated / sGDML potentials d <= 64,
                                 harmonic potentials (dense column pipeline with several CTAs per trajectory) d <= 96 */
#define SC_TIMING_SLOTS 8     /* per-kernel timing slots: path, rk4, lu, finish, rmult, potential Hessians */

typedef struct sc_potential sc_potential; /* opaque: device-resident potential parameters */
typedef struct sc_engine sc_engine;       /* opaque: one propagator instance (ensemble + constants) on one GPU */

int sc_abi_version(void);
const char *sc_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Potentials: batched energy / gradient / Hessian ("harmonic_approximation") and constant NAC vectors.
 * All constructor arguments are HOST pointers; parameters are copied to the current CUDA device.
 * --------------------------------------------------------------------------------------------- */

/* MorsePotential(omega, chi, nac): potentials.py:229-327.  a = sqrt(2 omega chi), D = omega/(4 chi) are
 * computed by the caller exactly as potentials.py:243-255 does (including the chi==0 -> 1e-4 bump);
 * all_harmonic selects the pure-harmonic branch (potentials.py:267-272).  masses == 1. */
int sc_potential_create_morse(sc_potential **out, int d, const double *omega_host, const double *a_host,
                              const double *D_host, int all_harmonic, const double *nac_host);

/* Morse potential in rotated coordinates x = Q r (dense Hessian Q h Q^T, dense-path fixture of SURVEY 8c-vi);
 * nac_host is the NAC vector in the ROTATED frame (Q tau1). */
int sc_potential_create_rotated_morse(sc_potential **out, int d, const double *omega_host, const double *a_host,
                                      const double *D_host, int all_harmonic, const double *nac_host,
                                      const double *Q_host);

/* NonHarmonicPotential(eps, b): potentials.py:25-205 (1-D Herman-Kluk test potential per mode, tau1 = 1) */
int sc_potential_create_nonharmonic(sc_potential **out, int d, const double *eps_host, const double *b_host);

/* MolecularHarmonicPotential: potentials.py:529-638.  V = energy0 - origin + g0.dr + dr.H0.dr/2 */
int sc_potential_create_harmonic(sc_potential **out, int d, const double *pos0_host, double energy0,
                                 const double *grad0_host, const double *hess0_host, const double *masses_host,
                                 const double *nac_host);

/* MolecularGDMLPotential / GDMLPredict: potentials.py:641-744, gdml_predictor.py:35-250.
 * xs_train, jx_alphas: (n_train_expanded, n_desc) row-major, already expanded over permutations
 * (gdml_predictor.py:67-82). */
int sc_potential_create_gdml(sc_potential **out, int n_atoms, int n_train, int n_desc, const double *xs_train_host,
                             const double *jx_alphas_host, double sig, double c, double std,
                             const double *masses_host, const double *nac_host);

/* energy origin subtracted from V (set by minimize(): potentials.py:523-526, 593, 699) */
int sc_potential_set_origin(sc_potential *pot, double origin);
int sc_potential_dimensions(const sc_potential *pot);
int sc_potential_destroy(sc_potential *pot);

/* potential.harmonic_approximation(r): r_dev (d, n) -> V_dev (n), grad_dev (d, n), hess_dev (d, d, n).
 * grad_dev / hess_dev may be NULL (energy only / energy+gradient). */
int sc_potential_eval(const sc_potential *pot, int n, const double *r_dev, double *V_dev, double *grad_dev,
                      double *hess_dev, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Propagator engine
 * --------------------------------------------------------------------------------------------- */

/* Host-side constants derived from Gamma_i, Gamma_t, Gamma_0, q0, p0 exactly as the reference does in
 * HermanKlukPropagator.__init__/initial_conditions/_prepare (propagators.py:408-443, 493-531, 633-643)
 * and WaltonManolopoulosPropagator._prepare (propagators.py:1102-1130). All HOST pointers, copied. */
typedef struct {
  int d;                 /* degrees of freedom */
  int dr;                /* rank of Gamma_i + Gamma_0 (d' in SURVEY.md) */
  int wm;                /* 0 = Herman-Kluk, 1 = Walton-Manolopoulos */
  /* prefactor factors with the subspace projection folded in (propagators.py:969-994):
   * L1 = U^T sqrt(Gamma_t), L2 = U^T Gamma_t^{-1/2}  (dr x d);  R1 = Gamma_i^{-1/2} U, R2 = sqrt(Gamma_i) U (d x dr) */
  const double *L1, *L2, *R1, *R2;
  const double *U;       /* d x dr, eigenvectors spanning the non-zero subspace (propagators.py:498) */
  const double *q0, *p0; /* centre of the initial wavepacket */
  /* coherent-state overlap constants (propagators.py:160-179, 230): A = Gi iGij Gj, B = iGij, C = Gj iGij, fac */
  const double *oi0_A, *oi0_B, *oi0_C; double oi0_fac;   /* <qi,pi,Gamma_i | q0,p0,Gamma_0> */
  const double *ot0_A, *ot0_B, *ot0_C; double ot0_fac;   /* <qt,pt,Gamma_t | q0,p0,Gamma_0> */
  const double *Gamma_0, *Gamma_i, *Gamma_t, *iGi0;       /* d x d */
  /* WM only */
  double alpha, beta;
  const double *iGamma_0;
  double detG0, detGi, detGt, detGi0;                     /* pi-absorbed pseudo-determinants (:1117-1125) */
} sc_engine_config;

int sc_engine_create(sc_engine **out, const sc_engine_config *cfg);
int sc_engine_destroy(sc_engine *eng);

/* initial_conditions, sampling part (propagators.py:533-555): x ~ N(0,1)^(2 d'), z = z0 + (Lz^-1)^T x, P = detLz/(2 pi)^d
 * exp(-x.x/2) on the device with a Philox4x32-10 counter-based generator -- zi_dev (2d, n) and probi_dev (n) are a pure
 * function of (seed, index0 + i), so ranks that pass the first global index of their shard draw ONE global ensemble.
 * iLq_host, iLp_host: the (d' x d) blocks of Lz^-1 (propagators.py:506-515), detLz (:531). */
int sc_engine_sample_ensemble(sc_engine *eng, int n, long long index0, unsigned long long seed, const double *iLq_host,
                              const double *iLp_host, double detLz, double *zi_dev, double *probi_dev, void *stream);
/* initial_conditions() after sampling (propagators.py:581-631): installs the ensemble (zi, probi), sets
 * Mqq = Mpp = 1, S = 0, t = 0, evaluates the prefactor once (initialises the sqrt branch trackers).
 * ntraj_norm is the N of the Monte-Carlo weight 1/(N probi (2 pi hbar)^d) -- the GLOBAL ensemble size when the
 * ensemble is sharded over ranks (propagators.py:837, 909). */
int sc_engine_set_ensemble(sc_engine *eng, int n, long long ntraj_norm, const double *zi_dev,
                           const double *probi_dev, void *stream);
int sc_engine_set_ensemble_host(sc_engine *eng, int n, long long ntraj_norm, const double *zi_host,
                                const double *probi_host, void *stream);

/* step(potential, dt) x nsteps fused with autocorrelation()/ic_correlation() of every new time
 * (propagators.py:645-655, 784-911; WM :1195-1389, 1577-1719).  For each of the nsteps new times writes
 *   corr_host[5*k + 0..1] = sum_n C_auto^(qp) / (N probi (2 pi)^d)      (no e^{i t E0} phase)
 *   corr_host[5*k + 2..3] = sum_n k_ic^(qp)   / (N probi (2 pi)^d) / hbar^2  (no phase)
 *   corr_host[5*k + 4]    = mean over the local ensemble of (T+V) at the 4th RK4 stage (propagators.py:380)
 * corr may be NULL.  The call synchronises the stream only when corr_host is given. */
int sc_engine_step(sc_engine *eng, const sc_potential *pot, double dt, int nsteps, double *corr_host,
                   void *stream);
/* same, results left on the device (nsteps x 5 doubles), no synchronisation */
int sc_engine_step_dev(sc_engine *eng, const sc_potential *pot, double dt, int nsteps, double *corr_dev,
                       void *stream);

/* autocorrelation() / ic_correlation(potential) at the CURRENT time without stepping; out_host[0..3] as above */
int sc_engine_correlations(sc_engine *eng, const sc_potential *pot, double *out_host, void *stream);

/* generic-potential stage interface (any Python object implementing the potential protocol):
 * one RK4 step = 4 x { sc_engine_stage_positions -> user evaluates V, grad, hess -> sc_engine_stage_apply },
 * then sc_engine_stage_finish (prefactor + branch tracking).  Arrays are batch-last device arrays. */
int sc_engine_stage_positions(sc_engine *eng, int stage, double dt, double *q_dev, void *stream);
int sc_engine_stage_apply(sc_engine *eng, int stage, double dt, const double *masses_dev, const double *V_dev,
                          const double *grad_dev, const double *hess_dev, double *energy_sum_dev, void *stream);
int sc_engine_stage_finish(sc_engine *eng, double dt, void *stream);
/* correlations with caller-supplied NAC data: n1 = -tau1/m as constant vector (d) on the host */
int sc_engine_correlations_n1(sc_engine *eng, const double *n1_host, double *out_host, void *stream);
/* the same two sums for potentials whose couplings depend on the position (propagators.py:868-909 in full generality;
 * Herman-Kluk): n1Q_dev, n1q_dev (d, n) = -hbar^2 tau1 / m at the current and at the initial positions
 * (potential.derivative_coupling_1st), n2Q_dev, n2q_dev (n) = -hbar^2/2 sum_k tau2_k / m_k (derivative_coupling_2nd) */
int sc_engine_correlations_general(sc_engine *eng, const double *n1Q_dev, const double *n1q_dev, const double *n2Q_dev,
                                   const double *n2q_dev, double *out_host, void *stream);

/* accessors (propagators.py:914-948): state in the reference's layout, prefactor and branch signs */
int sc_engine_get_state(sc_engine *eng, double *y_dev, void *stream);            /* (2d+4d^2+1, n) */
int sc_engine_set_state(sc_engine *eng, const double *y_dev, void *stream);
int sc_engine_get_prefactor(sc_engine *eng, double *c_dev /* c128 (n) sqrt(det), principal branch */,
                            double *c2_dev /* c128 (n) det */, double *signs_dev /* (3, n): C, detA, detM */,
                            void *stream);
/* ---- wavefunction diagnostics (Herman-Kluk): replace HermanKlukPropagator.coefficients / norm / wavefunction
 *      (propagators.py:657-686, 734-782, 688-732).  norm(): A, B, C = Gi (Gi+Gj)^+ Gj, (Gi+Gj)^+, Gj (Gi+Gj)^+ and fac of
 *      CoherentStatesOverlap(Gamma_t, Gamma_t) (propagators.py:174-179, 230), host (d x d) row-major; the result is the
 *      complex double sum (its real part is |psi|^2).  wavefunction(): x_dev (d, nx), fac = (det Gt / pi^rank)^(1/4). */
int sc_engine_coefficients(sc_engine *eng, double *v_dev /* c128 (n) */, void *stream);
int sc_engine_norm(sc_engine *eng, const double *A_host, const double *B_host, const double *C_host, double fac,
                   double *norm2_host /* [2] */, void *stream);
/* norm() of an ensemble that is sharded over ranks (propagators.py:734-782 is all pairs of the GLOBAL ensemble; cli.py:424-429):
 *   sc_engine_norm_pack   writes the ket vectors of this rank's shard into pack_dev (sc_engine_norm_pack_size doubles for
 *                         n_pad >= n_local rows; layout [r | s | alphaJ | beta | coef]) and keeps the bra side in the engine
 *   (caller)              all-gathers the packs of all ranks (NCCL)
 *   sc_engine_norm_block  adds fac * sum_{i local} conj(v_i) sum_{j in pack} <i|j> v_j to norm2_host[0:2] for ONE rank's pack
 *                         (n_ket = that rank's shard size); bra_coef_dev = coefficient section of this rank's own pack
 *   (caller)              sums over the packs, all-reduces the two doubles over the ranks, takes sqrt(Re) */
int sc_engine_norm_pack_size(const sc_engine *eng, int n_pad, long long *doubles_out);
int sc_engine_norm_pack(sc_engine *eng, const double *A_host, const double *B_host, const double *C_host, int n_pad,
                        double *pack_dev, void *stream);
int sc_engine_norm_block(sc_engine *eng, int n_ket, int n_pad, const double *pack_dev, const double *bra_coef_dev, double fac,
                         double *norm2_host, void *stream);
int sc_engine_wavefunction(sc_engine *eng, const double *Gamma_t_host, double fac, int nx, const double *x_dev,
                           double *phi_dev /* c128 (nx) */, void *stream);
int sc_engine_num_trajectories(const sc_engine *eng);
/* number of kernel launches issued by this engine so far (bench.py's gpu_launches) */
long long sc_engine_launch_count(const sc_engine *eng);
/* per-kernel device time of the column-chunked path (CUDA events on the launching stream): enable, run
 * sc_engine_step*, then read the accumulated milliseconds of { (q,p) path kernel, RK4/monodromy kernel, batched LU,
 * branch tracking + contributions } -- bench.py's roofline figure of the dominant kernel */
int sc_engine_set_timing(sc_engine *eng, int on);
int sc_engine_get_timing(sc_engine *eng, double *ms4_host);
/* all SC_TIMING_SLOTS slots: { path (+ overlap terms), RK4/monodromy, LU, finish, right factors (k_rmult), potential
 * Hessians, 0, 0 } of the column pipelines (propagators.py:645-655 has no counterpart: instrumentation only) */
int sc_engine_get_timing_slots(sc_engine *eng, double *ms_host, int nslots);
/* FP64 roofline denominator measured on the current device (no reference counterpart): register-resident chains of
 * mma.sync.m8n8k4.f64 and of DFMA on every SM, best of `reps`, CUDA events.  out_host[0] = DMMA TFLOP/s, [1] = DFMA TFLOP/s */
int sc_measure_fp64_peak(double *out_host, int reps, void *stream);
/* run-time options (no reference counterpart; measurement / diagnostics):
 *   "dense_engine" = 1  separable potentials (Morse / AS, NonHarmonic) are propagated by the general dense column pipeline
 *                       (sc_stream.cuh: their diagonal Hessians are expanded to full d x d matrices per stage) instead of
 *                       the structured pipeline of sc_chunk.cuh -- the configuration roofline figures of the dense engine
 *                       on the AS model are quoted on */
int sc_engine_set_option(sc_engine *eng, const char *name, int value);
/* name of the fused kernel variant the last sc_engine_step dispatched to (diagnostics) */
const char *sc_engine_kernel_name(const sc_engine *eng);

#ifdef __cplusplus
}
#endif
#endif /* SEMICLASSICAL_B200_H */
