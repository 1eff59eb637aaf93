"""
Named workloads (BASELINE.json configs): model parameters and time grids shared by the tests, the golden
generator (oracle/make_golden.py) and bench.py.  Pure numpy, no device code.

C1/C2: anharmonic adiabatic-shift (AS) model, 5 active modes of the reference fixture
       tests/DATA/AnharmonicAS/5modes/AS_model_chi0.0{0,2}.dat (values restated below) with the time grid of
       tests/test_propagators.py:378-382.
C4   : synthetic 60-mode AS model, generator fixed here.
"""
import numpy as np

from semiclassical_b200 import units

# omega / cm^-1, signed Huang-Rhys factor, non-adiabatic coupling  (5 uncommented rows of the fixture)
_AS5_ROWS = np.array([
    [500.8809000000, +0.3474950080, -0.0000460805],
    [827.3282000000, +0.3824004553, +0.0000595520],
    [990.0261000000, -0.4168571687, -0.0000150425],
    [1351.1072000000, -0.0935664944, +0.0002054889],
    [3256.3099000000, +0.0033317953, +0.0000665122],
])


class ASModel(object):
    """adiabatic-shift model: per-mode Morse (chi>0) or harmonic (chi==0) ground state, displaced
    harmonic excited state; see cli.py:229-285 for how the reference builds it from a model file"""
    def __init__(self, omega_cm, huang_rhys, nac, chi):
        omega_cm = np.asarray(omega_cm, dtype=np.float64)
        S = np.asarray(huang_rhys, dtype=np.float64)
        self.omega = omega_cm / units.hartree_to_wavenumbers
        self.nac = np.asarray(nac, dtype=np.float64).copy()
        self.chi = np.asarray(chi, dtype=np.float64).copy()
        # dQ = sqrt(2|S|/omega) sign(S)
        self.q0 = np.sqrt(2.0 * np.abs(S) / self.omega) * np.sign(S)
        self.p0 = np.zeros_like(self.q0)
        self.Gamma_0 = np.diag(self.omega)
        self.en_zpt = float(np.sum(0.5 * units.hbar * self.omega))
        self.dim = len(self.omega)


def as_5modes(chi=0.02):
    return ASModel(_AS5_ROWS[:, 0], _AS5_ROWS[:, 1], _AS5_ROWS[:, 2], np.full(5, chi))


def as_synthetic(dim=60, chi=0.02, seed=1234):
    """synthetic AS model of SURVEY 8d-C4: omega = linspace(200,3400) cm^-1, S = +-(0.02+0.08 U),
    nac = 1e-3 (U-0.5)"""
    rng = np.random.default_rng(seed)
    omega_cm = np.linspace(200.0, 3400.0, dim)
    S = (0.02 + 0.08 * rng.random(dim)) * np.where(rng.random(dim) < 0.5, -1.0, 1.0)
    nac = 1.0e-3 * (rng.random(dim) - 0.5)
    return ASModel(omega_cm, S, nac, np.full(dim, chi))


def test_time_grid():
    """(dt, nt) of tests/test_propagators.py:378-382: 100 points on [0, 150 fs / 40]"""
    nt = 4000 // 40
    t_max = 150.0 / units.autime_to_fs / 40.0
    times = np.linspace(0.0, t_max, nt)
    return float(times[1] - times[0]), nt


def random_orthogonal(dim, seed=7):
    """seeded dense orthogonal matrix (QR of a normal matrix, sign-fixed) for the rotated-AS fixtures"""
    rng = np.random.default_rng(seed)
    Q, R = np.linalg.qr(rng.standard_normal((dim, dim)))
    return Q * np.sign(np.diag(R))[None, :]


def gdml_synthetic(n_atoms=17, n_train=200, sig=80, seed=99, alpha_rms=2.0e3):
    """
    random-init sGDML model with the shapes of the coumarin fixture (N=17, D=136, M=200, identity
    permutation) plus a matching equilibrium-like geometry in bohr.  The training descriptors are
    inverse distances of jittered copies of the geometry so that the Matern kernel arguments have the
    magnitudes of a fitted model.
    """
    rng = np.random.default_rng(seed)
    # compact random molecule: points on a jittered grid, nearest distances ~ 2.6 bohr
    side = int(np.ceil(n_atoms ** (1.0 / 3.0)))
    grid = np.array([[i, j, k] for i in range(side) for j in range(side) for k in range(side)], dtype=float)
    pos = 2.6 * grid[:n_atoms] + 0.3 * rng.standard_normal((n_atoms, 3))
    i, j = np.tril_indices(n_atoms, -1)
    D = len(i)
    R_desc = np.zeros((D, n_train))
    for m in range(n_train):
        x = pos + 0.08 * rng.standard_normal((n_atoms, 3))
        R_desc[:, m] = 1.0 / np.linalg.norm(x[i] - x[j], axis=1)
    R_d_desc_alpha = alpha_rms * rng.standard_normal((n_train, D))
    model = {
        'sig': sig, 'c': -3.0, 'std': 0.05,
        'z': np.full(n_atoms, 6), 'R_desc': R_desc, 'R_d_desc_alpha': R_d_desc_alpha,
        'perms': np.arange(n_atoms)[None, :], 'tril_perms_lin': np.arange(D),
    }
    return model, pos.reshape(-1)


def harmonic_molecule_synthetic(dim=60, nzero=6, seed=3):
    """
    synthetic harmonic 'molecule' with the shape of the reference's molecular use case (C3 at size dim): dense
    Hessian with `nzero` zero modes (translations / rotations), unit masses, width matrix Gamma_0 = sqrt(Hessian)
    (rank dim - nzero, dense), wavepacket displaced along the vibrations only.
    Returns dict(pos0, hess0, grad0, energy0, masses, nac, Gamma_0, q0, p0, en_zpt).
    """
    rng = np.random.default_rng(seed)
    V, _ = np.linalg.qr(rng.standard_normal((dim, dim)))
    w = np.concatenate((np.zeros(nzero), np.linspace(200.0, 3400.0, dim - nzero) / units.hartree_to_wavenumbers))
    G0 = (V * w[None, :]) @ V.T
    hess = (V * (w * w)[None, :]) @ V.T
    pos0 = rng.standard_normal(dim)
    q0 = pos0 + V[:, nzero:] @ (0.3 * rng.standard_normal(dim - nzero) / np.sqrt(w[nzero:]))
    return dict(pos0=pos0, hess0=0.5 * (hess + hess.T), grad0=np.zeros(dim), energy0=0.0, masses=np.ones(dim),
                nac=1.0e-3 * rng.standard_normal(dim), Gamma_0=0.5 * (G0 + G0.T), q0=q0, p0=np.zeros(dim),
                en_zpt=float(0.5 * w.sum()))
