#!/usr/bin/env python
"""
TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the unmodified reference classes
(/root/reference, via oracle/refrun.py) on seeded inputs.  Run in the build container only:

    python oracle/make_golden.py            # (re)writes every fixture
    python oracle/make_golden.py hk_as5     # only fixtures whose name starts with the given prefix

Every fixture stores the complete input (model, Gamma's, wavepacket, injected ensemble zi/probi, time grid)
and the reference outputs (per-step correlation functions, final prefactors, branch signs, final state of
the first few trajectories), so the parity tests need nothing but the file.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refrun  # noqa: E402
from semiclassical_b200 import workloads  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
NKEEP = 8   # trajectories whose full final state vector y is stored


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def _propagate(name, propagators, potential, model_fields, Gi, Gt, G0, q0, p0, ntraj, dt, nt, en0,
               kind="HK", alpha=None, beta=None, seed=0, nkeep=NKEEP):
    torch.manual_seed(seed)
    if kind == "WM":
        pr = propagators.WaltonManolopoulosPropagator(T(Gi), T(Gt), alpha, beta)
    else:
        pr = propagators.HermanKlukPropagator(T(Gi), T(Gt))
    pr.initial_conditions(T(q0), T(p0), T(G0), ntraj=ntraj)
    zi, probi = pr.zi.numpy().copy(), pr.probi.numpy().copy()
    auto, ic = refrun.run_reference(pr, potential, dt, nt, en0)
    out = dict(model_fields)
    out.update(kind=kind, Gamma_i=Gi, Gamma_t=Gt, Gamma_0=G0, q0=q0, p0=p0, dt=dt, nt=nt, energy0_es=en0,
               zi=zi, probi=probi, autocorrelation=auto, ic_correlation=ic,
               t_final=float(pr.t),
               y_final=pr.y.numpy()[:, :nkeep].copy(),
               c_final=pr.c.numpy().copy(),
               c2_final=pr.sign_trackers["prefactorC"]["previous"].numpy().copy(),
               signs_C=pr.sign_trackers["prefactorC"]["signs"].numpy().real.copy())
    if kind == "WM":
        out.update(alpha=float(alpha), beta=float(beta),
                   signs_detA=pr.sign_trackers["detA"]["signs"].numpy().real.copy(),
                   signs_detM=pr.sign_trackers["detM"]["signs"].numpy().real.copy(),
                   detA_final=pr.detA.numpy().copy(), detM_final=pr.detM.numpy().copy(),
                   gamma_final=pr.gamma.numpy().copy())
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    flips = int((out["signs_C"] < 0).sum())
    print(f"{name:28s} n={ntraj:5d} nt={nt:4d} C(0)={auto[0]:.6f} |k_ic(0)|={abs(ic[0]):.3e} "
          f"flipped(C)={flips}  {os.path.getsize(path)/1024:.0f} KB")


def morse_case(name, propagators, potentials, model, ntraj, kind="HK", rotate_seed=None, nt=None, **kw):
    dt, nt_default = workloads.test_time_grid()
    nt = nt or nt_default
    pot = potentials.MorsePotential(T(model.omega.copy()), T(model.chi.copy()), T(model.nac.copy()))
    G = np.diag(model.omega)
    q0, p0 = model.q0, model.p0
    fields = dict(potential="morse", omega=model.omega, chi=model.chi, nac=model.nac)
    if rotate_seed is not None:
        Q = workloads.random_orthogonal(model.dim, rotate_seed)
        pot = refrun.RotatedPotential(pot, T(Q))
        G = Q @ G @ Q.T
        G = 0.5 * (G + G.T)
        q0, p0 = Q @ q0, Q @ p0
        fields.update(potential="rotated_morse", Q=Q)
    _propagate(name, propagators, pot, fields, G, G, G, q0, p0, ntraj, dt, nt, model.en_zpt, kind=kind, **kw)


def nonharmonic_case(name, propagators, potentials, ntraj, kind="HK", **kw):
    # tests/test_propagators.py:121-135, 261-275
    nt = 4000 // 40
    t_max = (12.0 / 40) * 2.0 * np.pi
    times = np.linspace(0.0, t_max, nt)
    dt = float(times[1] - times[0])
    pot = potentials.NonHarmonicPotential()
    fields = dict(potential="nonharmonic", eps=np.array([0.975]), b=np.array([12.0 ** -0.5]))
    Gi = np.array([[5.0]])
    G0 = np.array([[1.0]])
    _propagate(name, propagators, pot, fields, Gi, Gi, G0, np.array([7.3]), np.array([0.0]), ntraj, dt, nt, 0.5,
               kind=kind, **kw)


def methylium_case(name, propagators, potentials, readers, units, ntraj, nt, kind="HK", **kw):
    # cli.py:179-202, 293-313 on tests/DATA/examples/methylium_AH
    ddir = os.path.join(refrun.REFERENCE_ROOT, "tests", "DATA", "examples", "methylium_AH")
    with open(os.path.join(ddir, "opt_freq_s0.fchk")) as f:
        freq = readers.FormattedCheckpointFile(f)
    with open(os.path.join(ddir, "opt_freq_s1.fchk")) as f:
        exc = readers.FormattedCheckpointFile(f)
    pot = potentials.MolecularHarmonicPotential(freq, exc)
    x0, G0, en_zpt = exc.vibrational_groundstate()
    pot.minimize(T(x0))
    fields = dict(potential="harmonic", pos0=pot.pos0.numpy(), energy0=pot.energy0.numpy(), grad0=pot.grad0.numpy(),
                  hess0=pot.hess0.numpy(), nac=pot.nac0.numpy(), masses=pot._masses.numpy(), origin=pot._origin)
    dt = 0.005 / units.autime_to_fs
    _propagate(name, propagators, pot, fields, G0, G0, G0, x0, np.zeros_like(x0), ntraj, dt, nt, en_zpt, kind=kind, **kw)


class PositionDependentNAC(object):
    """potential protocol wrapper whose couplings depend on the position: tau1(r) = nac (1 + 0.2 r), tau2(r) = 0.05 nac r
    (elementwise) -- exercises propagators.py:868-909 in full generality; no shipped potential does"""
    def __init__(self, inner, nac):
        self.inner, self.nac = inner, nac

    def dimensions(self): return self.inner.dimensions()
    def masses(self): return self.inner.masses()
    def harmonic_approximation(self, r): return self.inner.harmonic_approximation(r)
    def derivative_coupling_1st(self, r): return self.nac.to(r.device).unsqueeze(1) * (1.0 + 0.2 * r)
    def derivative_coupling_2nd(self, r): return 0.05 * self.nac.to(r.device).unsqueeze(1) * r


def posnac_case(name, propagators, potentials, ntraj=300, nt=30, rotate_seed=None):
    model = workloads.as_5modes(0.02)
    dt, _ = workloads.test_time_grid()
    inner = potentials.MorsePotential(T(model.omega.copy()), T(model.chi.copy()), T(model.nac.copy()))
    G = np.diag(model.omega)
    q0, p0 = model.q0, model.p0
    fields = dict(potential="morse", omega=model.omega, chi=model.chi, nac=model.nac, posnac=1)
    nac = T(model.nac.copy())
    if rotate_seed is not None:
        Q = workloads.random_orthogonal(model.dim, rotate_seed)
        inner = refrun.RotatedPotential(inner, T(Q))
        G = Q @ G @ Q.T
        G = 0.5 * (G + G.T)
        q0, p0 = Q @ q0, Q @ p0
        nac = T(Q @ model.nac)
        fields.update(potential="rotated_morse", Q=Q)
    _propagate(name, propagators, PositionDependentNAC(inner, nac), fields, G, G, G, q0, p0, ntraj, dt, nt, model.en_zpt, seed=9)


def rates_case(name):
    """k_IC(E) by the reference's rates.rate_from_correlation (rates.py:20-82) with its gaussian and lorentzian lineshapes
    (broadening.py) on the IC correlation function of the hk_as5_chi002 fixture; row f4 of SURVEY section 8"""
    import importlib
    rates = importlib.import_module("semiclassical.rates")
    broadening = importlib.import_module("semiclassical.broadening")
    units = importlib.import_module("semiclassical.units")
    g = np.load(os.path.join(GOLDEN, "hk_as5_chi002.npz"))
    nt = int(g['nt'])
    times = np.linspace(0.0, nt * float(g['dt']), nt)              # the driver's grid (cli.py:312-313)
    corr = g['ic_correlation']
    sigma = 0.01 / np.sqrt(2.0 * np.log(2.0)) / units.hartree_to_ev
    gamma = 1.0e-3 / units.hartree_to_ev
    e1, r1 = rates.rate_from_correlation(times, corr, broadening.gaussian(sigma))
    e2, r2 = rates.rate_from_correlation(times, corr, broadening.lorentzian(gamma))
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, times=times, correlation=corr, sigma=sigma, gamma=gamma, energies=e1, rate_gaussian=r1,
                        energies_l=e2, rate_lorentzian=r2)
    print(f"{name:28s} nt={nt} max|k(E)| gaussian {np.abs(r1).max():.3e} lorentzian {np.abs(r2).max():.3e} "
          f"{os.path.getsize(path)/1024:.0f} KB")


def c2_full_size_case(name, propagators, potentials, ntraj=10000, seed=2002):
    """BASELINE configs[1] at its full size: AS 5 modes (chi = 0.02), Walton-Manolopoulos alpha = beta = 500, 10^4 trajectories,
    time grid of tests/test_propagators.py:378-382.  The ensemble is NOT stored: it is drawn by oracle.sample_ensemble from a
    numpy PCG64 stream (seed in the fixture, checksum stored) and injected into the reference propagator (SURVEY 8c recipe);
    the fixture holds the reference's correlation functions, the three branch-sign vectors and the final determinants."""
    import oracle as oracle   # this script runs from inside oracle/: the sibling module
    m = workloads.as_5modes(0.02)
    G = np.diag(m.omega)
    dt, nt = workloads.test_time_grid()
    zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, ntraj, np.random.default_rng(seed))
    pot = potentials.MorsePotential(T(m.omega.copy()), T(m.chi.copy()), T(m.nac.copy()))
    torch.manual_seed(0)
    pr = propagators.WaltonManolopoulosPropagator(T(G), T(G), 500, 500)
    pr.initial_conditions(T(m.q0), T(m.p0), T(G), ntraj=ntraj)
    d = len(m.q0)
    pr.zi = T(zi)
    pr.probi = T(probi)
    pr.y[:2 * d, :] = T(zi)
    del pr.sign_trackers
    pr._prefactor()
    auto, ic = refrun.run_reference(pr, pot, dt, nt, m.en_zpt)
    st = pr.sign_trackers
    out = dict(potential="morse", omega=m.omega, chi=m.chi, nac=m.nac, kind="WM", alpha=500.0, beta=500.0, Gamma_i=G, Gamma_t=G,
               Gamma_0=G, q0=m.q0, p0=m.p0, dt=dt, nt=nt, energy0_es=m.en_zpt, ensemble_seed=seed, ntraj=ntraj,
               zi_checksum=np.array([zi.sum(), np.abs(zi).sum(), zi[0, 0], zi[-1, -1]]), probi_sum=probi.sum(),
               autocorrelation=auto, ic_correlation=ic, t_final=float(pr.t),
               signs_C=st["prefactorC"]["signs"].numpy().real.astype(np.int8),
               signs_detA=st["detA"]["signs"].numpy().real.astype(np.int8),
               signs_detM=st["detM"]["signs"].numpy().real.astype(np.int8),
               c_final_sum=np.array([pr.c.numpy().sum()]), detA_final_sum=np.array([pr.detA.numpy().sum()]),
               detM_final_sum=np.array([pr.detM.numpy().sum()]))
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name:28s} n={ntraj:5d} nt={nt:4d} C(0)={auto[0]:.6f} flips C/A/M = {(out['signs_C'] < 0).sum()}/"
          f"{(out['signs_detA'] < 0).sum()}/{(out['signs_detM'] < 0).sum()}  {os.path.getsize(path)/1024:.0f} KB")


def gdml_potential_case(name, gdml_predictor, model, pos, nbatch, seed, jitter=0.05, store_model=False):
    rng = np.random.default_rng(seed)
    pred = gdml_predictor.GDMLPredict(model)
    r = pos[None, :] + jitter * rng.standard_normal((nbatch, len(pos)))
    e, g, h = pred.forward(T(r))
    path = os.path.join(GOLDEN, name + ".npz")
    extra = gdml_model_fields(model) if store_model else {}
    np.savez_compressed(path, r=r, energy=e.numpy(), grad=g.numpy(), hess=h.numpy(), **extra)
    print(f"{name:28s} B={nbatch} E[0]={e[0].item():.8f} |hess|max={h.abs().max().item():.3e} "
          f"{os.path.getsize(path)/1024:.0f} KB")


def fit_small_gdml(n_atoms=4, n_train=24, sig=12, seed=5, k_spring=0.3, jitter=0.15):
    """
    tiny *fitted* sGDML model (so that the dynamics fixture runs on a bound surface): the alphas solve the
    descriptor-space gradient equations of a pairwise-spring potential at the training geometries.
    The linear map alphas -> dE/dx is read off gdml_predictor.py:193-194.
    """
    rng = np.random.default_rng(seed)
    pos = np.array([[0.0, 0.0, 0.0], [2.1, 0.1, 0.0], [-0.7, 2.0, 0.2], [-0.6, -0.9, 1.9]])[:n_atoms]
    i, j = np.tril_indices(n_atoms, -1)
    D = len(i)
    r0 = np.linalg.norm(pos[i] - pos[j], axis=1)
    q = np.sqrt(5.0) / sig
    X = np.zeros((n_train, D))
    G = np.zeros((n_train, D))
    for m in range(n_train):
        x = pos + jitter * rng.standard_normal((n_atoms, 3))
        r = np.linalg.norm(x[i] - x[j], axis=1)
        X[m] = 1.0 / r
        G[m] = k_spring * (r - r0) * (-r * r)          # dV/d(1/r)
    K = np.zeros((n_train, D, n_train, D))
    for b in range(n_train):
        for m in range(n_train):
            diff = X[b] - X[m]
            rr = np.linalg.norm(diff)
            ef = q ** 4 / 3.0 * np.exp(-q * rr)
            K[b, :, m, :] = ef * (1.0 + q * rr) / q ** 2 * np.eye(D) - ef * np.outer(diff, diff)
    K = K.reshape(n_train * D, n_train * D)
    alphas = np.linalg.solve(K + 1.0e-10 * np.trace(K) / len(K) * np.eye(len(K)), G.reshape(-1)).reshape(n_train, D)
    model = {'sig': sig, 'c': 0.0, 'std': 1.0, 'z': np.array([6, 1, 7, 8])[:n_atoms], 'R_desc': X.T.copy(),
             'R_d_desc_alpha': alphas, 'perms': np.arange(n_atoms)[None, :], 'tril_perms_lin': np.arange(D)}
    return model, pos.reshape(-1)


def coumarin_model(units):
    """the reference's real sGDML fixture (tests/DATA/GDML: coumarin, 17 atoms, 200 training points, sig 80) and the
    geometry of coumarin.xyz in bohr; atomic masses from the element symbols"""
    ddir = os.path.join(refrun.REFERENCE_ROOT, "tests", "DATA", "GDML")
    model = dict(np.load(os.path.join(ddir, "coumarin_forces_au-wB97XD_def2SVP-train200-sym1.npz"), allow_pickle=True))
    xyz = np.loadtxt(os.path.join(ddir, "coumarin.xyz"), skiprows=2, usecols=(1, 2, 3)) / units.bohr_to_angs
    sym = np.loadtxt(os.path.join(ddir, "coumarin.xyz"), skiprows=2, usecols=(0,), dtype=str)
    amu = {"C": 12.011, "H": 1.008, "O": 15.999}
    masses = np.repeat(np.array([amu[s] for s in sym]) * units.amu_to_aumass, 3)
    return model, xyz.reshape(-1), masses


def gdml_model_fields(model):
    """the arrays a test needs to rebuild the model (they travel with the fixture: the reference's npz does not)"""
    return dict(gdml_sig=int(model['sig']), gdml_c=float(model['c']), gdml_std=float(model['std']),
                gdml_R_desc=np.asarray(model['R_desc'], dtype=np.float64),
                gdml_R_d_desc_alpha=np.asarray(model['R_d_desc_alpha'], dtype=np.float64))


def gdml_dynamics_case(name, propagators, potentials, gdml_predictor, ntraj, nt, model=None, pos=None, masses=None,
                       dt_fs=0.05, minimize=True, model_fixture=None, nkeep=NKEEP):
    """HK dynamics on an sGDML surface: default = small fitted model (N=4 atoms, d=12, Gamma of rank 6 like a real
    molecule); with model/pos/masses given: the real coumarin fixture (d=51, d'=45)"""
    if model is None:
        model, pos = fit_small_gdml()
        masses = np.repeat(np.array([12.0, 1.0, 14.0, 16.0]) * 1822.888486192, 3)
    d = len(pos)

    class _Nac(object):
        def __init__(self, d, z):
            self._d, self._z = d, z
        def nonadiabatic_coupling(self): return 1.0e-2 * np.cos(np.arange(self._d) + 1.0)
        def atomic_numbers(self): return self._z
        def masses(self): return masses
    nacf = _Nac(d, model['z'])
    pot = potentials.MolecularGDMLPotential(model, nacf)
    # energy origin at the minimum of the fitted surface (cli.py:293-295)
    if minimize:
        pot.minimize(T(pos))
    # widths from the Hessian at the (displaced) start geometry; the wavepacket then moves on the surface
    e, g, h = pot.harmonic_approximation(T(pos).unsqueeze(1))
    hm = h[:, :, 0].numpy() / np.sqrt(np.outer(masses, masses))
    w2, V = np.linalg.eigh(0.5 * (hm + hm.T))
    keep = w2 > 1.0e-7
    if d > 12:
        keep = np.zeros(d, dtype=bool)
        keep[6:] = True                      # eigh sorts ascending: the 6 smallest are translations / rotations
        assert w2[6] > 1.0e-7, w2[:8]
    L = np.sqrt(masses)[:, None] * V[:, keep] * (w2[keep] ** 0.25)[None, :]
    G0 = L @ L.T
    G0 = 0.5 * (G0 + G0.T)
    en0 = float(0.5 * np.sqrt(w2[keep]).sum())
    print("   gdml4: vib. frequencies (cm-1)", np.sqrt(w2[keep]) * 219474.63, " rank", keep.sum(), "origin", pot._origin)
    fields = dict(potential="gdml", nac=nacf.nonadiabatic_coupling(), masses=masses, origin=pot._origin)
    if model_fixture is None:
        fields.update(gdml_model_fields(model))
    else:
        fields.update(gdml_model_fixture=model_fixture)      # the model arrays live in that fixture
    dt = dt_fs / 0.02418884326505
    _propagate(name, propagators, pot, fields, G0, G0, G0, pos, np.zeros(d), ntraj, dt, nt, en0, nkeep=nkeep)


def diag_case(name, propagators, potential, fields, Gi, Gt, G0, q0, p0, ntraj, dt, nt, en0, nx, seed=0, xspread=0.3,
              kind="HK", alpha=None, beta=None):
    """wavefunction diagnostics of the reference HK propagator after nt steps: coefficients(), norm(), wavefunction(x)
    (propagators.py:657-782).  Stores q, p, S of every trajectory (not the monodromy blocks) besides the ensemble."""
    torch.manual_seed(seed)
    if kind == "WM":
        pr = propagators.WaltonManolopoulosPropagator(T(Gi), T(Gt), alpha, beta)
    else:
        pr = propagators.HermanKlukPropagator(T(Gi), T(Gt))
    pr.initial_conditions(T(q0), T(p0), T(G0), ntraj=ntraj)
    zi, probi = pr.zi.numpy().copy(), pr.probi.numpy().copy()
    auto, ic = refrun.run_reference(pr, potential, dt, nt, en0)
    d = len(q0)
    rng = np.random.default_rng(1000 + seed)
    qm = pr.y.numpy()[:d].mean(axis=1)
    x = qm[:, None] + xspread * rng.standard_normal((d, nx)) / np.sqrt(np.maximum(np.diag(Gt), 1e-3))[:, None]
    out = dict(fields)
    y = pr.y.numpy()
    if kind == "WM":
        out.update(alpha=float(alpha), beta=float(beta))
    out.update(kind=kind, Gamma_i=Gi, Gamma_t=Gt, Gamma_0=G0, q0=q0, p0=p0, dt=dt, nt=nt, energy0_es=en0, zi=zi, probi=probi,
               autocorrelation=auto, ic_correlation=ic, t_final=float(pr.t),
               qpS_final=np.concatenate((y[:2 * d], y[-1:]), axis=0).copy(), c_final=pr.c.numpy().copy(),
               signs_C=pr.sign_trackers["prefactorC"]["signs"].numpy().real.copy(),
               coefficients=pr.coefficients().numpy().copy(), norm=float(pr.norm()), x=x,
               wavefunction=np.asarray(pr.wavefunction(T(x))).copy())
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name:28s} n={ntraj:5d} nt={nt:4d} norm={out['norm']:.6f} max|psi|={np.abs(out['wavefunction']).max():.3e} "
          f"{os.path.getsize(path)/1024:.0f} KB")


def diag_methylium_wm(name, propagators, potentials, readers, units, ntraj, nt, nx, alpha, **kw):
    """WM diagnostics on the methylium harmonic model (d = 12, rank-deficient widths d' = 6): projection onto the non-zero
    subspace inside the all-pairs norm"""
    ddir = os.path.join(refrun.REFERENCE_ROOT, "tests", "DATA", "examples", "methylium_AH")
    with open(os.path.join(ddir, "opt_freq_s0.fchk")) as f:
        freq = readers.FormattedCheckpointFile(f)
    with open(os.path.join(ddir, "opt_freq_s1.fchk")) as f:
        exc = readers.FormattedCheckpointFile(f)
    pot = potentials.MolecularHarmonicPotential(freq, exc)
    x0, G0, en_zpt = exc.vibrational_groundstate()
    pot.minimize(T(x0))
    fields = dict(potential="harmonic", pos0=pot.pos0.numpy(), energy0=pot.energy0.numpy(), grad0=pot.grad0.numpy(),
                  hess0=pot.hess0.numpy(), nac=pot.nac0.numpy(), masses=pot._masses.numpy(), origin=pot._origin)
    dt = 0.005 / units.autime_to_fs
    diag_case(name, propagators, pot, fields, G0, G0, G0, x0, np.zeros_like(x0), ntraj, dt, nt, en_zpt, nx, kind="WM",
              alpha=alpha, beta=alpha, **kw)


def diag_morse(name, propagators, potentials, model, ntraj, nt, nx, rotate_seed=None, **kw):
    dt, _ = workloads.test_time_grid()
    pot = potentials.MorsePotential(T(model.omega.copy()), T(model.chi.copy()), T(model.nac.copy()))
    G = np.diag(model.omega)
    q0, p0 = model.q0, model.p0
    fields = dict(potential="morse", omega=model.omega, chi=model.chi, nac=model.nac)
    if rotate_seed is not None:
        Q = workloads.random_orthogonal(model.dim, rotate_seed)
        pot = refrun.RotatedPotential(pot, T(Q))
        G = Q @ G @ Q.T
        G = 0.5 * (G + G.T)
        q0, p0 = Q @ q0, Q @ p0
        fields.update(potential="rotated_morse", Q=Q)
    diag_case(name, propagators, pot, fields, G, G, G, q0, p0, ntraj, dt, nt, model.en_zpt, nx, **kw)


def main():
    prefixes = sys.argv[1:]
    os.makedirs(GOLDEN, exist_ok=True)
    propagators, potentials, units, readers, gdml_predictor = refrun.load_reference()

    def want(name):
        return (not prefixes) or any(name.startswith(p) for p in prefixes)

    if want("hk_as5_chi002"):
        morse_case("hk_as5_chi002", propagators, potentials, workloads.as_5modes(0.02), 1000)
    if want("hk_as5_chi000"):
        morse_case("hk_as5_chi000", propagators, potentials, workloads.as_5modes(0.0), 500)
    if want("wm_as5_chi002"):
        morse_case("wm_as5_chi002", propagators, potentials, workloads.as_5modes(0.02), 1000, kind="WM",
                   alpha=500, beta=500)
    if want("hk_as5_rot"):
        morse_case("hk_as5_rot", propagators, potentials, workloads.as_5modes(0.02), 500, rotate_seed=7)
    if want("wm_as5_rot"):
        morse_case("wm_as5_rot", propagators, potentials, workloads.as_5modes(0.02), 300, kind="WM",
                   alpha=500, beta=500, rotate_seed=7)
    if want("hk_1d"):
        nonharmonic_case("hk_1d", propagators, potentials, 2000)
    if want("wm_1d"):
        nonharmonic_case("wm_1d", propagators, potentials, 2000, kind="WM", alpha=100.0, beta=100.0)
    if want("hk_methylium"):
        methylium_case("hk_methylium", propagators, potentials, readers, units, 500, 200)
    if want("wm_methylium"):
        methylium_case("wm_methylium", propagators, potentials, readers, units, 200, 100, kind="WM",
                       alpha=1.0e4, beta=1.0e4)
    if want("hk_as60"):
        morse_case("hk_as60", propagators, potentials, workloads.as_synthetic(60), 64, nt=25)
    if want("hk_as60_rot"):
        morse_case("hk_as60_rot", propagators, potentials, workloads.as_synthetic(60), 48, nt=25, rotate_seed=11)
    if want("hk_as24_rot"):
        morse_case("hk_as24_rot", propagators, potentials, workloads.as_synthetic(24, seed=3), 96, nt=40, rotate_seed=5)
    if want("gdml_pot_n17"):
        model, pos = workloads.gdml_synthetic()
        gdml_potential_case("gdml_pot_n17", gdml_predictor, model, pos, 6, seed=1)
    if want("gdml_pot_n5"):
        model, pos = workloads.gdml_synthetic(n_atoms=5, n_train=16, sig=10, seed=3)
        gdml_potential_case("gdml_pot_n5", gdml_predictor, model, pos, 16, seed=2)
    if want("gdml_pot_coumarin"):
        # the reference's real fixture; the fitted model's arrays are stored with the outputs so that the tests can rebuild
        # the potential on the GPU box (the reference's npz does not travel)
        model, xyz, _ = coumarin_model(units)
        gdml_potential_case("gdml_pot_coumarin", gdml_predictor, model, xyz, 4, seed=4, jitter=0.02, store_model=True)
    if want("diag_as5"):
        diag_morse("diag_as5", propagators, potentials, workloads.as_5modes(0.02), 300, 30, 40)
    if want("diag_as5_rot"):
        diag_morse("diag_as5_rot", propagators, potentials, workloads.as_5modes(0.02), 150, 20, 33, rotate_seed=7, seed=2)
    if want("diag_as24"):
        diag_morse("diag_as24", propagators, potentials, workloads.as_synthetic(24, seed=3), 130, 10, 70, seed=3)
    if want("diag_1d"):
        nt = 50
        times = np.linspace(0.0, (12.0 / 40) * 2.0 * np.pi, 100)
        pot = potentials.NonHarmonicPotential()
        fields = dict(potential="nonharmonic", eps=np.array([0.975]), b=np.array([12.0 ** -0.5]))
        Gi = np.array([[5.0]])
        diag_case("diag_1d", propagators, pot, fields, Gi, Gi, np.array([[1.0]]), np.array([7.3]), np.array([0.0]), 500,
                  float(times[1] - times[0]), nt, 0.5, 128, seed=4, xspread=3.0)
    if want("hk_as5_posnac"):
        posnac_case("hk_as5_posnac", propagators, potentials)
    if want("hk_as5_rot_posnac"):
        posnac_case("hk_as5_rot_posnac", propagators, potentials, ntraj=200, nt=20, rotate_seed=7)
    if want("rates_as5"):
        rates_case("rates_as5")
    if want("c2_wm_as5_n10000"):
        c2_full_size_case("c2_wm_as5_n10000", propagators, potentials)
    if want("hk_gdml_coumarin"):
        # C5 at fixture size: HK dynamics on the real coumarin sGDML surface, d = 51, d' = 45, 24 steps
        model, xyz, masses = coumarin_model(units)
        gdml_dynamics_case("hk_gdml_coumarin", propagators, potentials, gdml_predictor, 40, 24, model=model, pos=xyz,
                           masses=masses, dt_fs=0.05, model_fixture="gdml_pot_coumarin", nkeep=3)
    if want("diag_wm_as5"):
        diag_morse("diag_wm_as5", propagators, potentials, workloads.as_5modes(0.02), 200, 20, 40, seed=5, kind="WM", alpha=500, beta=500)
    if want("diag_wm_as5_rot"):
        diag_morse("diag_wm_as5_rot", propagators, potentials, workloads.as_5modes(0.02), 120, 15, 33, rotate_seed=7, seed=6,
                   kind="WM", alpha=500, beta=500)
    if want("diag_wm_1d"):
        nt = 30
        times = np.linspace(0.0, (12.0 / 40) * 2.0 * np.pi, 100)
        pot = potentials.NonHarmonicPotential()
        fields = dict(potential="nonharmonic", eps=np.array([0.975]), b=np.array([12.0 ** -0.5]))
        Gi = np.array([[5.0]])
        diag_case("diag_wm_1d", propagators, pot, fields, Gi, Gi, np.array([[1.0]]), np.array([7.3]), np.array([0.0]), 300,
                  float(times[1] - times[0]), nt, 0.5, 64, seed=7, xspread=3.0, kind="WM", alpha=100.0, beta=100.0)
    if want("diag_wm_methylium"):
        diag_methylium_wm("diag_wm_methylium", propagators, potentials, readers, units, 90, 12, 25, 1.0e4, seed=8)
    if want("hk_gdml4"):
        gdml_dynamics_case("hk_gdml4", propagators, potentials, gdml_predictor, 200, 40)


if __name__ == "__main__":
    main()
