"""
GPU parity tests (run with -m gpu on the B200): the CUDA path, called through the C ABI via the Python mirror of
the reference interface, against (a) the golden vectors of the unmodified reference and (b) the C oracle on
seeded inputs.  Tolerance: max_t |C - C_ref| / max_t |C_ref| <= 1e-9 (BASELINE.json north_star), identical
sqrt-branch sign vectors.
"""
import numpy as np
import pytest
import torch

import helpers
from helpers import T, relerr

pytestmark = pytest.mark.gpu

TOL = 1.0e-9
HK_GOLDENS = ["hk_as5_chi002", "hk_as5_chi000", "hk_1d", "hk_as5_rot", "hk_methylium", "hk_as24_rot", "hk_as60",
              "hk_as60_rot"]


def run_loop(pr, pot, dt, nt, e0):
    """the reference's driver loop (cli.py:401-436)"""
    auto, ic = np.zeros(nt, complex), np.zeros(nt, complex)
    for k in range(nt):
        auto[k] = pr.autocorrelation(e0)
        ic[k] = pr.ic_correlation(pot, e0)
        pr.step(pot, dt)
    return auto, ic


@pytest.mark.parametrize("name", HK_GOLDENS)
def test_hk_step_loop_matches_reference(name, cuda_device):
    g = helpers.load_golden(name)
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, cuda_device)
    auto, ic = run_loop(pr, pot, float(g['dt']), int(g['nt']), float(g['energy0_es']))
    assert relerr(auto, g['autocorrelation']) < TOL
    assert relerr(ic, g['ic_correlation']) < TOL
    nk = g['y_final'].shape[1]
    assert relerr(pr.y[:, :nk].cpu().numpy(), g['y_final']) < TOL
    assert relerr(pr.c.cpu().numpy(), g['c_final']) < TOL
    signs = pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy()
    assert np.array_equal(signs, g['signs_C'])
    assert abs(float(pr.t) - float(g['t_final'])) < 1e-12 * max(1.0, abs(float(g['t_final'])))


@pytest.mark.parametrize("name", HK_GOLDENS)
def test_hk_fused_propagate_matches_reference(name, cuda_device):
    g = helpers.load_golden(name)
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, cuda_device)
    nt, e0 = int(g['nt']), float(g['energy0_es'])
    a0, i0 = pr.autocorrelation(e0), pr.ic_correlation(pot, e0)
    a, i = pr.propagate(pot, float(g['dt']), nt - 1, e0)
    assert relerr(np.concatenate(([a0], a)), g['autocorrelation']) < TOL
    assert relerr(np.concatenate(([i0], i)), g['ic_correlation']) < TOL


@pytest.mark.parametrize("name", ["hk_as5_chi002", "hk_methylium"])
def test_per_trajectory_contributions_sum_to_autocorrelation(name, cuda_device):
    """autocorrelation_qp() (propagators.py:784-807): the Monte-Carlo sum of the per-trajectory contributions is the
    autocorrelation function the step kernels accumulate"""
    g = helpers.load_golden(name)
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, cuda_device)
    e0, dt = float(g['energy0_es']), float(g['dt'])
    for k in range(3):
        qp = pr.autocorrelation_qp()
        assert qp.shape == (pr.ntraj,)
        total = torch.sum(qp / (pr.ntraj * pr.probi * (2 * np.pi)**pr.dim)) * np.exp(1j * float(pr.t) * e0)
        assert abs(complex(total) - g['autocorrelation'][k]) < 1e-11 * max(1.0, abs(g['autocorrelation'][k]))
        assert abs(complex(total) - pr.autocorrelation(e0)) < 1e-11
        pr.step(pot, dt)


def test_initial_autocorrelation_is_one(cuda_device):
    g = helpers.load_golden("hk_as5_chi002")
    pr = helpers.propagator_from_golden(g, cuda_device)
    assert abs(pr.autocorrelation(float(g['energy0_es'])) - 1.0) < 1e-12


def test_shards_sum_to_whole(cuda_device):
    """multi-GPU decomposition on one device: sum over shards == unsharded"""
    g = helpers.load_golden("hk_as5_rot")
    pot = helpers.potential_from_golden(g)
    n, nt, e0, dt = len(g['probi']), 30, float(g['energy0_es']), float(g['dt'])
    acc_a, acc_i = np.zeros(nt, complex), np.zeros(nt, complex)
    for sl in (slice(0, 137), slice(137, 300), slice(300, n)):
        pr = helpers.propagator_from_golden(g, cuda_device, nslice=sl)
        a, i = pr.propagate(pot, dt, nt, e0)
        acc_a += a
        acc_i += i
    assert relerr(acc_a, g['autocorrelation'][1:nt + 1]) < TOL
    assert relerr(acc_i, g['ic_correlation'][1:nt + 1]) < TOL


def test_engine_against_oracle_on_fresh_ensemble(cuda_device):
    """seeded ensemble that is in no fixture: CUDA path vs the C oracle, 5-mode AS, 4096 trajectories"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.as_5modes(0.02)
    G = np.diag(m.omega)
    zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, 4096, np.random.default_rng(42))
    dt, nt = workloads.test_time_grid()
    ref = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), oracle.Consts(G, G, G, m.q0, m.p0), zi, probi,
                     dt, nt, m.en_zpt)
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi), T(probi))
    auto, ic = run_loop(pr, pot, dt, nt, m.en_zpt)
    assert relerr(auto, ref['autocorrelation']) < TOL
    assert relerr(ic, ref['ic_correlation']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), ref['signs'][0])


@pytest.mark.parametrize("name", ["hk_as5_chi002", "hk_1d", "hk_methylium", "hk_as5_rot"])
def test_potential_kernels_against_oracle(name, cuda_device):
    """batched harmonic_approximation kernels vs the oracle's potentials"""
    from oracle import oracle
    g = helpers.load_golden(name)
    pot = helpers.potential_from_golden(g)
    opot, _, _ = oracle.from_golden(g)
    d = pot.dimensions()
    r = g['zi'][:d, :257].copy()
    V, grad, hess = pot.harmonic_approximation(T(r).to(cuda_device))
    Vo, go, ho = opot.eval(r)
    assert np.abs(V.cpu().numpy() - Vo).max() <= 1e-12 * max(1.0, np.abs(Vo).max())
    assert relerr(grad.cpu().numpy(), go) < 1e-12
    assert relerr(hess.cpu().numpy(), ho) < 1e-12


def test_energy_guard_raises(cuda_device):
    """a time step far too large violates energy conservation -> RuntimeError (propagators.py:385-398)"""
    g = helpers.load_golden("hk_1d")
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, cuda_device)
    with pytest.raises(RuntimeError, match="not conserved"):
        for _ in range(50):
            pr.step(pot, 3.0)


def test_wrong_dimension_is_rejected(cuda_device):
    g = helpers.load_golden("hk_1d")
    pr = helpers.propagator_from_golden(g, cuda_device)
    pot5 = helpers.potential_from_golden(helpers.load_golden("hk_as5_chi002"))
    with pytest.raises(AssertionError):
        pr.step(pot5, 0.1)


# ------------------------------------------------------------------ Walton-Manolopoulos (config 2)
WM_GOLDENS = ["wm_1d", "wm_as5_chi002", "wm_as5_rot", "wm_methylium"]


def _check_wm_signs(pr, g):
    st = pr.sign_trackers
    assert np.array_equal(st["prefactorC"]["signs"].real.cpu().numpy(), g['signs_C'])
    assert np.array_equal(st["detA"]["signs"].real.cpu().numpy(), g['signs_detA'].real)
    assert np.array_equal(st["detM"]["signs"].real.cpu().numpy(), g['signs_detM'].real)


@pytest.mark.parametrize("name", WM_GOLDENS)
def test_wm_step_loop_matches_reference(name, cuda_device):
    g = helpers.load_golden(name)
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, cuda_device)
    auto, ic = run_loop(pr, pot, float(g['dt']), int(g['nt']), float(g['energy0_es']))
    assert relerr(auto, g['autocorrelation']) < TOL
    assert relerr(ic, g['ic_correlation']) < TOL
    nk = g['y_final'].shape[1]
    assert relerr(pr.y[:, :nk].cpu().numpy(), g['y_final']) < TOL
    _check_wm_signs(pr, g)


@pytest.mark.parametrize("name", WM_GOLDENS)
def test_wm_fused_propagate_matches_reference(name, cuda_device):
    g = helpers.load_golden(name)
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, cuda_device)
    nt, e0 = int(g['nt']), float(g['energy0_es'])
    a0, i0 = pr.autocorrelation(e0), pr.ic_correlation(pot, e0)
    a, i = pr.propagate(pot, float(g['dt']), nt - 1, e0)
    assert relerr(np.concatenate(([a0], a)), g['autocorrelation']) < TOL
    assert relerr(np.concatenate(([i0], i)), g['ic_correlation']) < TOL
    pr.step(pot, float(g['dt']))
    _check_wm_signs(pr, g)


@pytest.mark.parametrize("d,rotated", [(14, False), (16, True), (24, False), (24, True), (32, True), (60, False)])
def test_wm_at_larger_dimensions(d, rotated, cuda_device):
    """Walton-Manolopoulos beyond the small systems: a CTA per trajectory with the workspace in shared memory up to 29 modes
    (full-rank widths: 16 d^2 complex numbers), in global-memory slabs above (k_wm_global).  AS model with d modes
    (optionally rotated: dense Hessians and dense widths) against the C oracle, three branch trackers bit-identical."""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.as_synthetic(d, 0.02)
    if rotated:
        Q = workloads.random_orthogonal(d)
        G = Q @ np.diag(m.omega) @ Q.T
        G = 0.5 * (G + G.T)
        q0, p0 = Q @ m.q0, Q @ m.p0
        opot = oracle.Potential.rotated_morse(m.omega, m.chi, m.nac, Q)
        pot = potentials.RotatedMorsePotential(T(m.omega), T(m.chi), T(m.nac), T(Q))
    else:
        G, q0, p0 = np.diag(m.omega), m.q0, m.p0
        opot = oracle.Potential.morse(m.omega, m.chi, m.nac)
        pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    n, nt = 40, 6
    zi, probi = oracle.sample_ensemble(G, G, q0, p0, n, np.random.default_rng(5))
    dt, _ = workloads.test_time_grid()
    ref = oracle.run(opot, oracle.Consts(G, G, G, q0, p0, alpha=500.0, beta=500.0), zi, probi, dt, nt, m.en_zpt, wm=True)
    pr = propagators.WaltonManolopoulosPropagator(T(G), T(G), 500, 500, device=cuda_device)
    pr.set_ensemble(T(q0), T(p0), T(G), T(zi), T(probi))
    auto, ic = run_loop(pr, pot, dt, nt, m.en_zpt)
    assert pr.kernel_name().endswith("k_wm_global" if d > 29 else "+k_wm" if d > 16 else "+k_wm_fused")
    assert relerr(auto, ref['autocorrelation']) < TOL
    assert relerr(ic, ref['ic_correlation']) < TOL
    st = pr.sign_trackers
    assert np.array_equal(st["prefactorC"]["signs"].real.cpu().numpy(), ref['signs'][0])
    assert np.array_equal(st["detA"]["signs"].real.cpu().numpy(), ref['signs'][1])
    assert np.array_equal(st["detM"]["signs"].real.cpu().numpy(), ref['signs'][2])


def test_c2_full_size_wm_fused_launches(cuda_device):
    """BASELINE configs[1] at its FULL size: AS 5 modes, WM alpha = beta = 500, 10^4 trajectories, 100 steps, against the
    reference's golden (ensemble regenerated from its numpy seed).  K-step fused launches (k_hk_generic snapshots +
    ONE k_wm_fused per 33 steps); all three branch-sign vectors identical"""
    from semiclassical_b200 import propagators
    g = helpers.load_golden("c2_wm_as5_n10000")
    zi, probi = helpers.regenerate_ensemble(g)
    pot = helpers.potential_from_golden(g)
    pr = propagators.WaltonManolopoulosPropagator(T(g['Gamma_i']), T(g['Gamma_t']), 500, 500, device=cuda_device)
    pr.set_ensemble(T(g['q0']), T(g['p0']), T(g['Gamma_0']), T(zi), T(probi))
    nt, e0, dt = int(g['nt']), float(g['energy0_es']), float(g['dt'])
    auto, ic = [pr.autocorrelation(e0)], [pr.ic_correlation(pot, e0)]
    l0 = pr.launch_count()
    for k0 in range(0, nt - 1, 33):
        a, i = pr.propagate(pot, dt, min(33, nt - 1 - k0), e0)
        auto.extend(a)
        ic.extend(i)
    assert pr.kernel_name() == "k_hk_small+k_wm_fused"
    assert pr.launch_count() - l0 <= 3 * 5 + 3          # launches per 33 fused steps, not per step
    assert relerr(auto, g['autocorrelation']) < TOL
    assert relerr(ic, g['ic_correlation']) < TOL
    pr.step(pot, dt)
    _check_wm_signs(pr, g)


# ------------------------------------------------------------------ sGDML (config 5)
# the kernel sum over training points is cancellation-prone: the reference's own gradient / Hessian move by
# ~3e-9 (absolute) when the training set is summed in another order; the kernel's
# summation order differs from torch's, hence the looser potential-level tolerance.
GDML_TOL = 1.0e-8


@pytest.mark.parametrize("name,kwargs", [("gdml_pot_n17", {}), ("gdml_pot_n5", dict(n_atoms=5, n_train=16, sig=10, seed=3))])
def test_gdml_kernel_matches_reference(name, kwargs, cuda_device):
    from semiclassical_b200 import workloads, potentials
    model, pos = workloads.gdml_synthetic(**kwargs)
    g = helpers.load_golden(name)
    d = len(pos)
    pot = potentials.MolecularGDMLPotential.from_arrays(model, np.ones(d), np.zeros(d))
    r = T(g['r'].T.copy()).to(cuda_device)
    V, grad, hess = pot.harmonic_approximation(r)
    V, grad, hess = V.cpu().numpy(), grad.cpu().numpy(), hess.cpu().numpy()
    assert np.abs(V - g['energy']).max() < GDML_TOL * max(1.0, np.abs(g['energy']).max())
    assert relerr(grad.T, g['grad']) < GDML_TOL
    assert relerr(hess.transpose(2, 0, 1), g['hess']) < GDML_TOL
    assert np.abs(hess - hess.transpose(1, 0, 2)).max() < 1.0e-13 * np.abs(hess).max()


def _coumarin_model(g):
    d = g['r'].shape[1]
    return dict(sig=int(g['gdml_sig']), c=float(g['gdml_c']), std=float(g['gdml_std']), R_desc=g['gdml_R_desc'],
                R_d_desc_alpha=g['gdml_R_d_desc_alpha'], perms=np.arange(d // 3)[None, :],
                tril_perms_lin=np.arange(g['gdml_R_desc'].shape[0])), d


def test_gdml_kernel_on_the_real_coumarin_fixture(cuda_device):
    """k_gdml_eval on the reference's real fitted model (coumarin, 17 atoms, 200 training points, sig 80) vs the
    reference's own GDMLPredict.forward outputs (potential-level parity on the real fixture)"""
    from semiclassical_b200 import potentials
    g = helpers.load_golden("gdml_pot_coumarin")
    model, d = _coumarin_model(g)
    pot = potentials.MolecularGDMLPotential.from_arrays(model, np.ones(d), np.zeros(d))
    V, grad, hess = pot.harmonic_approximation(T(g['r'].T.copy()).to(cuda_device))
    V, grad, hess = V.cpu().numpy(), grad.cpu().numpy(), hess.cpu().numpy()
    # fitted alphas ~7e11: 12 digits cancel in the kernel sum; the reference's own outputs move by 2.3e-9 / 3.8e-9 / 3.2e-9
    # absolute (E / grad / Hessian) under a permutation of its training set -- the bar is a few times that
    assert np.abs(V - g['energy']).max() < 2.0e-8
    assert np.abs(grad.T - g['grad']).max() < 2.0e-8 and relerr(grad.T, g['grad']) < 5.0e-7
    assert np.abs(hess.transpose(2, 0, 1) - g['hess']).max() < 2.0e-8 and relerr(hess.transpose(2, 0, 1), g['hess']) < 5.0e-8
    assert np.abs(hess - hess.transpose(1, 0, 2)).max() < 1.0e-13 * np.abs(hess).max()


def test_hk_on_the_real_coumarin_sgdml_surface(cuda_device):
    """C5 at fixture size: HK dynamics (d = 51, d' = 45, 40 trajectories, 24 steps) on the real coumarin sGDML surface vs the
    reference's golden -- fused launches on the dense column pipeline (k_gdml_eval writes the Hessian of every RK4 stage
    straight into the stream image k_rk4_stream consumes) and, second, step by step through the drop-in API.
    Tolerance 1e-7: the kernel sum over the training set is cancellation-prone (the reference's own gradient moves by
    ~4e-9 under a permutation of the training points) and 24 steps of dynamics amplify it."""
    g = helpers.load_golden("hk_gdml_coumarin")
    pot = helpers.potential_from_golden(g)
    nt, e0, dt = int(g['nt']), float(g['energy0_es']), float(g['dt'])
    pr = helpers.propagator_from_golden(g, cuda_device)
    a0, i0 = pr.autocorrelation(e0), pr.ic_correlation(pot, e0)
    a, i = pr.propagate(pot, dt, nt - 1, e0)
    assert pr.kernel_name().startswith("k_rk4_stream+k_rmult+")
    assert relerr(np.concatenate(([a0], a)), g['autocorrelation']) < 1.0e-7
    assert relerr(np.concatenate(([i0], i)), g['ic_correlation']) < 1.0e-7
    pr.step(pot, dt)
    nk = g['y_final'].shape[1]
    assert relerr(pr.y[:, :nk].cpu().numpy(), g['y_final']) < 1.0e-7
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), g['signs_C'])
    pr2 = helpers.propagator_from_golden(g, cuda_device)
    auto, ic = run_loop(pr2, pot, dt, nt, e0)
    assert relerr(auto, g['autocorrelation']) < 1.0e-7
    assert relerr(ic, g['ic_correlation']) < 1.0e-7


def test_gdml_kernel_large_batch_against_oracle(cuda_device):
    """1000 geometries (several per CTA, every tile phase exercised) vs the C oracle"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials
    model, pos = workloads.gdml_synthetic()
    d = len(pos)
    rng = np.random.default_rng(5)
    r = pos[:, None] + 0.05 * rng.standard_normal((d, 1000))
    opot = oracle.Potential.gdml(model, np.ones(d), np.zeros(d))
    Vo, go, ho = opot.eval(r)
    pot = potentials.MolecularGDMLPotential.from_arrays(model, np.ones(d), np.zeros(d))
    V, grad, hess = pot.harmonic_approximation(T(r).to(cuda_device))
    assert np.abs(V.cpu().numpy() - Vo).max() < GDML_TOL * max(1.0, np.abs(Vo).max())
    assert relerr(grad.cpu().numpy(), go) < GDML_TOL
    assert relerr(hess.cpu().numpy(), ho) < GDML_TOL


def test_hk_with_gdml_potential_matches_reference(cuda_device):
    """HK dynamics driven by the sGDML kernel through the stage interface (4-atom fixture, 200 trajectories)"""
    g = helpers.load_golden("hk_gdml4")
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, cuda_device)
    auto, ic = run_loop(pr, pot, float(g['dt']), int(g['nt']), float(g['energy0_es']))
    assert relerr(auto, g['autocorrelation']) < 1.0e-8
    assert relerr(ic, g['ic_correlation']) < 1.0e-8
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), g['signs_C'])


# ------------------------------------------------------------------ generic-potential stage interface
class _PythonPotential(object):
    """duck-typed potential (no native handle): the protocol of SURVEY.md section 8b"""
    def __init__(self, inner):
        self.inner = inner

    def dimensions(self):
        return self.inner.dimensions()

    def masses(self):
        return self.inner.masses()

    def harmonic_approximation(self, r):
        return self.inner.harmonic_approximation(r)

    def derivative_coupling_1st(self, r):
        return self.inner.derivative_coupling_1st(r)

    def derivative_coupling_2nd(self, r):
        return self.inner.derivative_coupling_2nd(r)


class _PositionDependentNAC(_PythonPotential):
    """tau1(r) = nac (1 + 0.2 r), tau2(r) = 0.05 nac r -- the wrapper the fixtures hk_as5*_posnac were generated with"""
    def __init__(self, inner, nac):
        super().__init__(inner)
        self.nac = nac

    def derivative_coupling_1st(self, r):
        return self.nac.to(r.device).unsqueeze(1) * (1.0 + 0.2 * r)

    def derivative_coupling_2nd(self, r):
        return 0.05 * self.nac.to(r.device).unsqueeze(1) * r


@pytest.mark.parametrize("name", ["hk_as5_posnac", "hk_as5_rot_posnac"])
def test_position_dependent_couplings_match_reference(name, cuda_device):
    """ic_correlation with couplings that depend on the position (propagators.py:868-909 in full generality, no shipped
    potential has them): the potential object supplies tau1, tau2 at the initial and current positions, k_corr_general the
    rest; diagonal and dense (rotated) width matrices, against the reference's golden"""
    g = helpers.load_golden(name)
    inner = helpers.potential_from_golden(g)
    nac = T(g['Q'] @ g['nac']) if 'Q' in g.files else T(g['nac'])
    pot = _PositionDependentNAC(inner, nac)
    pr = helpers.propagator_from_golden(g, cuda_device)
    nt = int(g['nt'])
    auto, ic = run_loop(pr, pot, float(g['dt']), nt, float(g['energy0_es']))
    assert relerr(auto, g['autocorrelation']) < TOL
    assert relerr(ic, g['ic_correlation']) < TOL


@pytest.mark.parametrize("name", ["hk_as5_chi002", "hk_methylium", "hk_as24_rot", "wm_as5_rot", "hk_as60", "hk_as60_rot"])
def test_python_potential_through_stage_interface(name, cuda_device):
    """any object with the potential protocol drives the RK4 stages; d >= 17 runs on the dense column pipeline (the
    caller's (d, d, n) Hessians become the stream images of the four stages), smaller d on the DFMA stage kernel"""
    g = helpers.load_golden(name)
    pot = _PythonPotential(helpers.potential_from_golden(g))
    pr = helpers.propagator_from_golden(g, cuda_device)
    nt = min(int(g['nt']), 40)
    auto, ic = run_loop(pr, pot, float(g['dt']), nt, float(g['energy0_es']))
    assert relerr(auto, g['autocorrelation'][:nt]) < TOL
    assert relerr(ic, g['ic_correlation'][:nt]) < TOL
    if nt == int(g['nt']):
        nk = g['y_final'].shape[1]
        assert relerr(pr.y[:, :nk].cpu().numpy(), g['y_final']) < TOL
        assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), g['signs_C'])


def test_python_potential_at_d72(cuda_device):
    """a user potential object for a 24-atom harmonic molecule (d = 72, d' = 66: beyond the DFMA stage kernel's shared memory)
    through the stage interface on the dense column pipeline, vs the C oracle"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    d = 72
    m = workloads.harmonic_molecule_synthetic(d)
    G = m['Gamma_0']
    n, nt = 37, 5
    zi, probi = oracle.sample_ensemble(G, G, m['q0'], m['p0'], n, np.random.default_rng(72))
    dt, _ = workloads.test_time_grid()
    opot = oracle.Potential.harmonic(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
    ref = oracle.run(opot, oracle.Consts(G, G, G, m['q0'], m['p0']), zi, probi, dt, nt, m['en_zpt'])
    pot = _PythonPotential(potentials.MolecularHarmonicPotential.from_arrays(m['pos0'], m['energy0'], m['grad0'], m['hess0'],
                                                                             m['masses'], m['nac']))
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m['q0']), T(m['p0']), T(G), T(zi), T(probi))
    auto, ic = run_loop(pr, pot, dt, nt, m['en_zpt'])
    assert pr.kernel_name().startswith("stage interface: k_rk4_stream")
    assert relerr(auto, ref['autocorrelation']) < TOL
    assert relerr(ic, ref['ic_correlation']) < TOL
    assert relerr(pr.y.cpu().numpy(), ref['y']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), ref['signs'][0])


# ------------------------------------------------------------------ column-chunked headline path
@pytest.mark.parametrize("d", [17, 18, 23, 24, 25, 32, 33, 40, 51, 60, 62, 63, 64])
def test_chunked_path_against_oracle(d, cuda_device):
    """diagonal-Gamma AS models on the chunked RK4 + batched-LU path (sc_chunk.cuh) vs the C oracle: ragged batch
    (n not a multiple of anything), two launches of KC steps, odd chunk widths (d = 51), 4 chunks (d = 62); from d = 17
    (the smallest instantiation) through both LU kernels (d <= 24: k_lu_warp)"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.as_synthetic(d)
    G = np.diag(m.omega)
    n, nt = 301, 13
    zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, n, np.random.default_rng(100 + d))
    dt, _ = workloads.test_time_grid()
    ref = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), oracle.Consts(G, G, G, m.q0, m.p0), zi, probi,
                     dt, nt, m.en_zpt)
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi), T(probi))
    a0, i0 = pr.autocorrelation(m.en_zpt), pr.ic_correlation(pot, m.en_zpt)
    a, i = pr.propagate(pot, dt, nt - 1, m.en_zpt)
    assert pr.kernel_name().startswith("k_rk4_wcols+")
    assert relerr(np.concatenate(([a0], a)), ref['autocorrelation']) < TOL
    assert relerr(np.concatenate(([i0], i)), ref['ic_correlation']) < TOL
    # one more step through the drop-in API, then state / prefactor / branch signs of all trajectories
    pr.step(pot, dt)
    ref2 = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), oracle.Consts(G, G, G, m.q0, m.p0), zi, probi,
                      dt, nt, m.en_zpt)
    assert relerr(pr.y.cpu().numpy(), ref2['y']) < TOL
    assert relerr(pr.c.cpu().numpy(), ref2['c']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), ref2['signs'][0])


# ------------------------------------------------------------------ wavefunction diagnostics (propagators.py:657-782)
@pytest.mark.parametrize("name", ["diag_as5", "diag_as5_rot", "diag_as24", "diag_1d", "diag_wm_as5", "diag_wm_as5_rot", "diag_wm_1d",
                                  "diag_wm_methylium"])
def test_wavefunction_diagnostics_match_reference(name, cuda_device):
    """coefficients(), norm() (all-pairs DMMA kernel), wavefunction(x) after nt steps vs the reference's own values; the
    diag_wm_* fixtures are the Walton-Manolopoulos versions (propagators.py:1391-1575: eqn (75) coefficients, per-pair
    (d' x d') complex inverse + determinant in the norm; methylium: rank-deficient widths, d' = 6 of d = 12)"""
    g = helpers.load_golden(name)
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, cuda_device)
    pr.propagate(pot, float(g['dt']), int(g['nt']), float(g['energy0_es']))
    assert relerr(pr.coefficients().cpu().numpy(), g['coefficients']) < TOL
    assert abs(pr.norm() / float(g['norm']) - 1.0) < TOL
    phi = pr.wavefunction(T(g['x']))
    assert phi.shape == g['wavefunction'].shape
    assert relerr(phi, g['wavefunction']) < TOL


def test_wm_diagnostics_against_oracle_on_fresh_ensemble(cuda_device):
    """Walton-Manolopoulos coefficients / wavefunction / norm on an ensemble that is in no fixture (n = 257: ragged for every
    kernel shape) vs the oracle's restatement of propagators.py:1391-1575"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.as_5modes(0.02)
    G = np.diag(m.omega)
    n, nt = 257, 7
    zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, n, np.random.default_rng(4711))
    dt, _ = workloads.test_time_grid()
    consts = oracle.Consts(G, G, G, m.q0, m.p0, 300.0, 300.0)
    ref = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), consts, zi, probi, dt, nt, m.en_zpt, wm=True)
    pc = oracle.wm_diag_pieces(consts, ref['y'], zi)
    v = oracle.wm_coefficients(consts, zi, probi, ref['y'], ref['c'], ref['signs'][0], ref['signs'][1], pc)
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    pr = propagators.WaltonManolopoulosPropagator(T(G), T(G), 300.0, 300.0, device=cuda_device)
    pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi), T(probi))
    pr.propagate(pot, dt, nt, m.en_zpt)
    assert relerr(pr.coefficients().cpu().numpy(), v) < TOL
    x = m.q0[:, None] + 0.05 * np.random.default_rng(6).standard_normal((5, 23))
    assert relerr(pr.wavefunction(T(x)), oracle.wm_wavefunction(consts, zi, ref['y'], v, pc, x)) < TOL
    assert abs(pr.norm() / oracle.wm_norm(consts, zi, ref['y'], v, pc) - 1.0) < TOL


def test_sharded_norm_blocks_add_up(cuda_device):
    """norm() of a sharded ensemble (propagators.py:734-782 is all pairs of the GLOBAL ensemble): two shards on one device,
    the four (n_a x n_b) blocks through sc_engine_norm_pack / sc_engine_norm_block add up to the reference's norm"""
    g = helpers.load_golden("diag_as24")
    pot = helpers.potential_from_golden(g)
    n = len(g['probi'])
    cut = 57
    shards = [helpers.propagator_from_golden(g, cuda_device, nslice=sl) for sl in (slice(0, cut), slice(cut, n))]
    for pr in shards:
        pr.propagate(pot, float(g['dt']), int(g['nt']), float(g['energy0_es']))
    n_pad = max(pr.ntraj for pr in shards)
    packs = [pr.norm_pack(n_pad) for pr in shards]
    tot = 0.0j
    for a, pa in enumerate(shards):
        for b, pb in enumerate(shards):
            tot += pa.norm_block(packs[b], pb.ntraj, n_pad, packs[a])
    assert abs(tot.imag) < 1.0e-12 * abs(tot.real)
    assert abs(np.sqrt(tot.real) / float(g['norm']) - 1.0) < TOL


def test_norm_against_oracle_ragged(cuda_device):
    """norm() on an ensemble that is not a multiple of the 64 x 32 tile (n = 203, d = 40: K = 80) vs the numpy oracle"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.as_synthetic(40)
    G = np.diag(m.omega)
    n, nt = 203, 5
    zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, n, np.random.default_rng(77))
    dt, _ = workloads.test_time_grid()
    ref = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), oracle.Consts(G, G, G, m.q0, m.p0), zi, probi, dt, nt,
                     m.en_zpt)
    d = 40
    qpS = np.concatenate((ref['y'][:2 * d], ref['y'][-1:]), axis=0)
    v = oracle.hk_coefficients(G, G, m.q0, m.p0, zi, probi, qpS, ref['c'], ref['signs'][0])
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi), T(probi))
    pr.propagate(pot, dt, nt, m.en_zpt)
    assert relerr(pr.coefficients().cpu().numpy(), v) < TOL
    assert abs(pr.norm() / oracle.hk_norm(G, qpS, v) - 1.0) < TOL
    x = m.q0[:, None] + 0.05 * np.random.default_rng(5).standard_normal((d, 19))
    assert relerr(pr.wavefunction(T(x)), oracle.hk_wavefunction(G, qpS, v, x)) < TOL


# ------------------------------------------------------------------ on-device ensemble sampling (propagators.py:493-555)
@pytest.mark.parametrize("name", ["hk_as5_chi002", "hk_methylium", "hk_as24_rot"])
def test_device_sampler(name, cuda_device):
    """k_sample_ensemble (Philox4x32-10 + Box-Muller) behind initial_conditions: the ensemble is a pure function of
    (seed, global index) -- shards concatenate bitwise to the unsharded draw; probi is exactly the density of the recovered
    normals (diagonal, rank-deficient d' = 6 of 12, and dense widths); moments and a Kolmogorov-Smirnov test of the normals"""
    from scipy import stats
    from semiclassical_b200 import propagators
    g = helpers.load_golden(name)
    pr = propagators.HermanKlukPropagator(T(g['Gamma_i']), T(g['Gamma_t']), device=cuda_device)
    q0, p0, G0 = T(g['q0']), T(g['p0']), T(g['Gamma_0'])
    n = 100000
    zi, probi = pr.sample_ensemble(q0, p0, G0, n, seed=77)
    za, pa = pr.sample_ensemble(q0, p0, G0, 40001, index0=0, seed=77)
    zb, pb = pr.sample_ensemble(q0, p0, G0, n - 40001, index0=40001, seed=77)
    assert torch.equal(torch.cat((za, zb), dim=1), zi) and torch.equal(torch.cat((pa, pb)), probi)
    z2, _ = pr.sample_ensemble(q0, p0, G0, 1000, seed=78)
    assert not torch.equal(z2, zi[:, :1000])
    # recover the normals: zi - z0 = (Lz^-1)^T x  (least squares: exact in the sampled subspace)
    iLq, iLp, detLz = pr._iLz_detLz
    d, dr = pr.dim, pr.rank
    iLz = np.zeros((2 * dr, 2 * d))
    iLz[:dr, :d], iLz[dr:, d:] = iLq.numpy(), iLp.numpy()
    dz = zi.cpu().numpy() - np.concatenate((g['q0'], g['p0']))[:, None]
    x, res, _, _ = np.linalg.lstsq(iLz.T, dz, rcond=None)
    assert np.abs(iLz.T @ x - dz).max() < 1e-10 * max(1.0, np.abs(dz).max())
    pref = detLz / (2 * np.pi) ** d
    assert np.abs(probi.cpu().numpy() / (pref * np.exp(-0.5 * (x * x).sum(axis=0))) - 1.0).max() < 1e-9
    # statistics of the 2 d' normals
    assert np.abs(x.mean(axis=1)).max() < 5.0 / np.sqrt(n)
    assert np.abs(np.cov(x) - np.eye(2 * dr)).max() < 6.0 / np.sqrt(n)
    for j in (0, dr - 1, dr, 2 * dr - 1):
        assert stats.kstest(x[j], 'norm').pvalue > 1e-4


# ------------------------------------------------------------------ `semi dynamics` driver (cli.py:171-476)
def test_dynamics_driver_matches_reference_loop(tmp_path, cuda_device):
    """the JSON-task driver on the 5-mode AS fixture with the reference's ensemble injected: same correlations.npz
    content as the reference loop (golden), running average over repetitions, times grid quirk, overwrite=false"""
    from semiclassical_b200 import dynamics, units, workloads
    g = helpers.load_golden("hk_as5_chi002")
    m = workloads.as_5modes(0.02)
    model_file = tmp_path / "AS_model.dat"
    rows = workloads._AS5_ROWS
    np.savetxt(model_file, np.column_stack((rows, np.full(len(rows), 0.02))), fmt="%.10f",
               header="omega/cm^-1  Huang-Rhys  NAC  chi")
    nt, n = int(g['nt']), len(g['probi'])
    out = tmp_path / "correlations.npz"
    task = {"task": "dynamics", "potential": {"type": "anharmonic AS", "model_file": str(model_file)}, "propagator": "HK",
            "batch_size": n, "num_trajectories": n, "num_steps": nt, "time_step_fs": float(g['dt']) * units.autime_to_fs,
            "results": {"correlations": str(out)}}
    dynamics.run_semiclassical_dynamics(task, device=cuda_device, ensembles=[(g['zi'], g['probi'])], steps_per_launch=17)
    data = dict(np.load(out))
    assert relerr(data['autocorrelation'], g['autocorrelation']) < TOL
    assert relerr(data['ic_correlation'], g['ic_correlation']) < TOL
    assert int(data['trajectories']) == n and str(data['propagator']) == "HK"
    assert np.allclose(data['times'], np.linspace(0.0, nt * float(g['dt']), nt), rtol=1e-12)
    assert abs(float(data['zero_point_energy']) - m.en_zpt) < 1e-12
    # second run accumulates into the same file: identical ensemble -> identical averages, doubled count
    task["results"]["overwrite"] = False
    dynamics.run_semiclassical_dynamics(task, device=cuda_device, ensembles=[(g['zi'], g['probi'])])
    data2 = dict(np.load(out))
    assert int(data2['trajectories']) == 2 * n
    assert relerr(data2['autocorrelation'], data['autocorrelation']) < 1.0e-13
    # two repetitions in one call (one propagator, buffers reused): running average of identical ensembles
    task["results"]["overwrite"] = True
    task["num_trajectories"] = 2 * n
    dynamics.run_semiclassical_dynamics(task, device=cuda_device, ensembles=[(g['zi'], g['probi'])] * 2, steps_per_launch=33)
    data4 = dict(np.load(out))
    assert int(data4['trajectories']) == 2 * n
    assert relerr(data4['autocorrelation'], g['autocorrelation']) < TOL and relerr(data4['ic_correlation'], g['ic_correlation']) < TOL
    task["num_trajectories"] = n
    # sampled on the device (torch CUDA generator): C(0) = 1 and the statistical agreement the reference tests ask for
    task["results"]["overwrite"] = True
    task["manual_seed"] = 0
    task["calc_norm_every"] = 50
    dynamics.run_semiclassical_dynamics(task, device=cuda_device)
    data3 = dict(np.load(out))
    assert abs(data3['autocorrelation'][0] - 1.0) < 1.0e-9
    assert relerr(data3['autocorrelation'], g['autocorrelation']) < 0.2


# ------------------------------------------------------------------ general path: dense Gamma, rank deficient
@pytest.mark.parametrize("d", [13, 16, 23, 30, 31, 38, 40, 54, 60, 64, 72, 96])
def test_dense_harmonic_molecule_against_oracle(d, cuda_device):
    """dense Hessian + dense width matrices with 6 zero modes (d' = d - 6) on the dense column pipeline (sc_stream.cuh):
    Hessian and left prefactor factors streamed through the shared-memory ring, k_rmult for the right factors (k-padding at
    d = 23), batched LU with d' = 7, 10, 17, 24 (k_lu_warp), 25, 32, 34 (k_lu_mma, ragged last panel), 48, 54, 58"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.harmonic_molecule_synthetic(d)
    G = m['Gamma_0']
    n, nt = 157, 9
    zi, probi = oracle.sample_ensemble(G, G, m['q0'], m['p0'], n, np.random.default_rng(200 + d))
    dt, _ = workloads.test_time_grid()
    opot = oracle.Potential.harmonic(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
    ref = oracle.run(opot, oracle.Consts(G, G, G, m['q0'], m['p0']), zi, probi, dt, nt, m['en_zpt'])
    pot = potentials.MolecularHarmonicPotential.from_arrays(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m['q0']), T(m['p0']), T(G), T(zi), T(probi))
    a0, i0 = pr.autocorrelation(m['en_zpt']), pr.ic_correlation(pot, m['en_zpt'])
    a, i = pr.propagate(pot, dt, nt - 1, m['en_zpt'])
    assert pr.kernel_name().startswith("k_rk4_stream+k_rmult+")
    assert relerr(np.concatenate(([a0], a)), ref['autocorrelation']) < TOL
    assert relerr(np.concatenate(([i0], i)), ref['ic_correlation']) < TOL
    pr.step(pot, dt)
    ref2 = oracle.run(opot, oracle.Consts(G, G, G, m['q0'], m['p0']), zi, probi, dt, nt, m['en_zpt'])
    assert relerr(pr.y.cpu().numpy(), ref2['y']) < TOL
    assert relerr(pr.c.cpu().numpy(), ref2['c']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), ref2['signs'][0])


@pytest.mark.parametrize("d,nzero", [(2, 0), (3, 1), (4, 2), (6, 0), (6, 3), (7, 1), (8, 2), (9, 3), (10, 4), (11, 0), (12, 5), (12, 0)])
def test_small_systems_any_rank(d, nzero, cuda_device):
    """k_hk_small with dense width matrices of rank d' = d - nzero <= d (the factors are zero-padded to d and the padded
    diagonal of the prefactor matrix set to one): dense harmonic 'molecules' with 2 ... 12 coordinates against the C oracle"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.harmonic_molecule_synthetic(d, nzero, seed=40 + d)
    G = m['Gamma_0']
    n, nt = 131, 12
    zi, probi = oracle.sample_ensemble(G, G, m['q0'], m['p0'], n, np.random.default_rng(300 + d))
    dt, _ = workloads.test_time_grid()
    opot = oracle.Potential.harmonic(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
    ref = oracle.run(opot, oracle.Consts(G, G, G, m['q0'], m['p0']), zi, probi, dt, nt, m['en_zpt'])
    pot = potentials.MolecularHarmonicPotential.from_arrays(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m['q0']), T(m['p0']), T(G), T(zi), T(probi))
    assert pr.rank == d - nzero
    a0, i0 = pr.autocorrelation(m['en_zpt']), pr.ic_correlation(pot, m['en_zpt'])
    a, i = pr.propagate(pot, dt, nt - 1, m['en_zpt'])
    assert pr.kernel_name() == "k_hk_small"
    assert relerr(np.concatenate(([a0], a)), ref['autocorrelation']) < TOL
    assert relerr(np.concatenate(([i0], i)), ref['ic_correlation']) < TOL
    pr.step(pot, dt)
    ref2 = oracle.run(opot, oracle.Consts(G, G, G, m['q0'], m['p0']), zi, probi, dt, nt, m['en_zpt'])
    assert relerr(pr.y.cpu().numpy(), ref2['y']) < TOL
    assert relerr(pr.c.cpu().numpy(), ref2['c']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), ref2['signs'][0])


@pytest.mark.parametrize("d,n,nt", [(1, 1, 5), (1, 17, 40), (5, 1, 6), (5, 2, 6), (5, 7, 130), (5, 1777, 4), (12, 1, 5), (12, 5, 70),
                                    (3, 11, 9), (9, 3, 9)])
def test_small_kernel_edge_shapes(d, n, nt, cuda_device):
    """k_hk_small on ragged ensembles: a single trajectory (one group of a warp active), fewer trajectories than groups per
    warp, ensembles that are not a multiple of the trajectories per warp / CTA, more steps per launch than any other test
    (129 fused steps), AS model with d modes against the C oracle"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.as_synthetic(d, 0.02) if d != 5 else workloads.as_5modes(0.02)
    G = np.diag(m.omega)
    zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, n, np.random.default_rng(900 + 10 * d + n))
    dt, _ = workloads.test_time_grid()
    ref = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), oracle.Consts(G, G, G, m.q0, m.p0), zi, probi, dt, nt, m.en_zpt)
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi), T(probi))
    a0, i0 = pr.autocorrelation(m.en_zpt), pr.ic_correlation(pot, m.en_zpt)
    a, i = pr.propagate(pot, dt, nt - 1, m.en_zpt)
    assert pr.kernel_name() == "k_hk_small"
    assert relerr(np.concatenate(([a0], a)), ref['autocorrelation']) < TOL
    assert relerr(np.concatenate(([i0], i)), ref['ic_correlation']) < TOL
    pr.step(pot, dt)
    assert relerr(pr.y.cpu().numpy(), ref['y']) < TOL
    assert relerr(pr.c.cpu().numpy(), ref['c']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), ref['signs'][0])


@pytest.mark.parametrize("d,n,nt", [(13, 5, 4), (14, 1, 3), (16, 150, 5), (17, 1, 4), (18, 2, 5), (31, 149, 4), (32, 3, 37), (33, 150, 3), (47, 297, 3), (61, 5, 4), (65, 7, 3),
                                    (80, 151, 3)])
def test_dense_pipeline_edge_shapes(d, n, nt, cuda_device):
    """ragged shapes of the dense column pipeline: smallest d, d not a multiple of 4 / 8, one or two trajectories (fewer than
    SMs, than warps of a path CTA), ensembles that are not a multiple of the SM count, more steps than one pass holds (37 > 16:
    three passes over the state), tile groups with a ragged last group (d = 65: 17 tiles over 8-warp CTAs)"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.harmonic_molecule_synthetic(d)
    G = m['Gamma_0']
    zi, probi = oracle.sample_ensemble(G, G, m['q0'], m['p0'], n, np.random.default_rng(500 + d))
    dt, _ = workloads.test_time_grid()
    opot = oracle.Potential.harmonic(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
    ref = oracle.run(opot, oracle.Consts(G, G, G, m['q0'], m['p0']), zi, probi, dt, nt + 1, m['en_zpt'])
    pot = potentials.MolecularHarmonicPotential.from_arrays(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m['q0']), T(m['p0']), T(G), T(zi), T(probi))
    a, i = pr.propagate(pot, dt, nt, m['en_zpt'])
    assert pr.kernel_name().startswith("k_rk4_stream+k_rmult+")
    assert relerr(a, ref['autocorrelation'][1:]) < TOL
    assert relerr(i, ref['ic_correlation'][1:]) < TOL
    pr.step(pot, dt)                                   # the oracle loop of nt + 1 reads ends with one more step
    assert relerr(pr.y.cpu().numpy(), ref['y']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), ref['signs'][0])


@pytest.mark.parametrize("d", [14, 20, 40])
def test_separable_potential_with_dense_widths(d, cuda_device):
    """AS (Morse) potential -- diagonal Hessian -- with DENSE width matrices (a wavepacket whose widths are not aligned with the
    modes): k_path_separable + k_aux_terms + expanded diagonal Hessians + the dense prefactor stages, vs the C oracle"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.as_synthetic(d, seed=77)
    Q = workloads.random_orthogonal(d, 21)
    G = Q @ np.diag(m.omega * (1.0 + 0.3 * np.cos(np.arange(d)))) @ Q.T
    G = 0.5 * (G + G.T)
    n, nt = 131, 9
    zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, n, np.random.default_rng(600 + d))
    dt, _ = workloads.test_time_grid()
    opot = oracle.Potential.morse(m.omega, m.chi, m.nac)
    ref = oracle.run(opot, oracle.Consts(G, G, G, m.q0, m.p0), zi, probi, dt, nt + 1, m.en_zpt)
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi), T(probi))
    a0, i0 = pr.autocorrelation(m.en_zpt), pr.ic_correlation(pot, m.en_zpt)
    a, i = pr.propagate(pot, dt, nt, m.en_zpt)
    assert pr.kernel_name().startswith("k_rk4_stream+k_rmult+")
    assert relerr(np.concatenate(([a0], a)), ref['autocorrelation']) < TOL
    assert relerr(np.concatenate(([i0], i)), ref['ic_correlation']) < TOL
    pr.step(pot, dt)
    assert relerr(pr.y.cpu().numpy(), ref['y']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), ref['signs'][0])


@pytest.mark.parametrize("name", ["hk_as24_rot", "hk_as60_rot"])
def test_rotated_models_run_on_the_stream_pipeline(name, cuda_device):
    """per-trajectory dense Hessians (Q diag(h) Q^T formed by k_expand_hessian) + dense width matrices: the reference's own
    goldens of the rotated AS models, on k_rk4_stream"""
    g = helpers.load_golden(name)
    pot = helpers.potential_from_golden(g)
    pr = helpers.propagator_from_golden(g, cuda_device)
    nt, e0 = int(g['nt']), float(g['energy0_es'])
    a0, i0 = pr.autocorrelation(e0), pr.ic_correlation(pot, e0)
    a, i = pr.propagate(pot, float(g['dt']), nt - 1, e0)
    assert pr.kernel_name().startswith("k_rk4_stream+k_rmult+")
    assert relerr(np.concatenate(([a0], a)), g['autocorrelation']) < TOL
    assert relerr(np.concatenate(([i0], i)), g['ic_correlation']) < TOL
    pr.step(pot, float(g['dt']))
    nk = g['y_final'].shape[1]
    assert relerr(pr.y[:, :nk].cpu().numpy(), g['y_final']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), g['signs_C'])


@pytest.mark.parametrize("d", [15, 20, 33, 60, 64])
def test_dense_engine_option_on_separable_model(d, cuda_device):
    """AS model through the general dense pipeline (diagonal stage Hessians expanded to full matrices, option
    'dense_engine') vs the C oracle -- the configuration the dense-engine roofline of configs[3] is quoted on"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    m = workloads.as_synthetic(d)
    G = np.diag(m.omega)
    n, nt = 211, 11
    zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, n, np.random.default_rng(300 + d))
    dt, _ = workloads.test_time_grid()
    ref = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), oracle.Consts(G, G, G, m.q0, m.p0), zi, probi,
                     dt, nt, m.en_zpt)
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi), T(probi))
    pr.set_option("dense_engine", 1)
    a0, i0 = pr.autocorrelation(m.en_zpt), pr.ic_correlation(pot, m.en_zpt)
    a, i = pr.propagate(pot, dt, nt - 1, m.en_zpt)
    assert pr.kernel_name().startswith("k_rk4_stream+k_lu")
    assert relerr(np.concatenate(([a0], a)), ref['autocorrelation']) < TOL
    assert relerr(np.concatenate(([i0], i)), ref['ic_correlation']) < TOL
    pr.step(pot, dt)
    ref2 = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), oracle.Consts(G, G, G, m.q0, m.p0), zi, probi,
                      dt, nt, m.en_zpt)
    assert relerr(pr.y.cpu().numpy(), ref2['y']) < TOL
    assert np.array_equal(pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy(), ref2['signs'][0])


@pytest.mark.parametrize("which", ["as_structured", "as_dense_engine", "harmonic_dense"])
def test_long_horizon_d60(which, cuda_device):
    """the reference's full time grid (100 steps of tests/test_propagators.py:378-382) at d = 60 on both column pipelines:
    the 1e-9 bar and the bit-identical branch signs must hold over the whole horizon, not only over the 25 steps of the
    hk_as60* fixtures (fused launches of 33 steps: several passes over the state, branch flips included)"""
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    dt, nt = workloads.test_time_grid()
    n = 40
    if which == "harmonic_dense":
        m = workloads.harmonic_molecule_synthetic(60)
        G, q0, p0, e0 = m['Gamma_0'], m['q0'], m['p0'], m['en_zpt']
        opot = oracle.Potential.harmonic(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
        pot = potentials.MolecularHarmonicPotential.from_arrays(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
    else:
        m = workloads.as_synthetic(60)
        G, q0, p0, e0 = np.diag(m.omega), m.q0, m.p0, m.en_zpt
        opot = oracle.Potential.morse(m.omega, m.chi, m.nac)
        pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    zi, probi = oracle.sample_ensemble(G, G, q0, p0, n, np.random.default_rng(900))
    ref = oracle.run(opot, oracle.Consts(G, G, G, q0, p0), zi, probi, dt, nt, e0)
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
    pr.set_ensemble(T(q0), T(p0), T(G), T(zi), T(probi))
    if which == "as_dense_engine":
        pr.set_option("dense_engine", 1)
    auto, ic = [pr.autocorrelation(e0)], [pr.ic_correlation(pot, e0)]
    for k0 in range(0, nt - 1, 33):
        a, i = pr.propagate(pot, dt, min(33, nt - 1 - k0), e0)
        auto.extend(a)
        ic.extend(i)
    expected = {"as_structured": "k_rk4_wcols", "as_dense_engine": "k_rk4_stream+k_lu", "harmonic_dense": "k_rk4_stream+k_rmult"}[which]
    assert pr.kernel_name().startswith(expected)
    assert relerr(auto, ref['autocorrelation']) < TOL
    assert relerr(ic, ref['ic_correlation']) < TOL
    pr.step(pot, dt)
    assert relerr(pr.y.cpu().numpy(), ref['y']) < TOL
    assert relerr(pr.c.cpu().numpy(), ref['c']) < TOL
    signs = pr.sign_trackers["prefactorC"]["signs"].real.cpu().numpy()
    assert np.array_equal(signs, ref['signs'][0])


# ------------------------------------------------------------------ BASELINE size: size-independent properties
def test_full_size_properties_c4(cuda_device):
    """configs[3] at its full single-GPU size (10^6 trajectories, 60 modes, ~116 GB of state) where no oracle run is
    affordable: C(0) = 1 exactly (Gamma_i = Gamma_0, cli.py:460-467), the correlation functions are linear in the
    ensemble (two half ensembles normalised by the global N add up to the whole), a repeated run is bitwise identical,
    and a 1 000-trajectory subset agrees with the C oracle"""
    import gc
    from oracle import oracle
    from semiclassical_b200 import workloads, potentials, propagators
    free, _ = torch.cuda.mem_get_info(torch.device(cuda_device))
    n = 1000000 if free > 150e9 else 200000
    m = workloads.as_synthetic(60)
    G = np.diag(m.omega)
    zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, n, np.random.default_rng(2024))
    dt, _ = workloads.test_time_grid()
    nt = 20
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))

    def run(sl, repeat=False):
        pr = propagators.HermanKlukPropagator(T(G), T(G), device=cuda_device)
        pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi[:, sl]), T(probi[sl]), ntraj_total=n)
        c0 = pr.autocorrelation(m.en_zpt)
        a, i = pr.propagate(pot, dt, nt, m.en_zpt)
        if repeat:
            pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi[:, sl]), T(probi[sl]), ntraj_total=n)
            a2, i2 = pr.propagate(pot, dt, nt, m.en_zpt)
            assert np.array_equal(a, a2) and np.array_equal(i, i2)
        del pr
        gc.collect()
        torch.cuda.empty_cache()
        return c0, a, i

    c0, a, i = run(slice(0, n), repeat=True)
    assert abs(c0 - 1.0) < 1.0e-9
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(i))
    h = n // 2 + 12345
    c0a, aa, ia = run(slice(0, h))
    c0b, ab, ib = run(slice(h, n))
    assert abs(c0a + c0b - c0) < 1.0e-12
    assert relerr(aa + ab, a) < 1.0e-12 and relerr(ia + ib, i) < 1.0e-12
    # subset vs oracle (normalised by the same global N)
    ns = 1000
    ref = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), oracle.Consts(G, G, G, m.q0, m.p0), zi[:, :ns].copy(),
                     probi[:ns].copy(), dt, nt + 1, m.en_zpt, ntraj_norm=n)
    _, asub, isub = run(slice(0, ns))
    assert relerr(asub, ref['autocorrelation'][1:]) < TOL and relerr(isub, ref['ic_correlation'][1:]) < TOL
