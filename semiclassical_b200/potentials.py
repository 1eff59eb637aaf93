"""
Potential energy surfaces with the interface of the reference's semiclassical/potentials.py
(dimensions / masses / harmonic_approximation / derivative_coupling_1st / derivative_coupling_2nd),
evaluated by sm_100a kernels through the C ABI.  Batch-last layout: r is (dim, n).

  NonHarmonicPotential        potentials.py:25-205
  MorsePotential              potentials.py:208-397   (the anharmonic adiabatic-shift model)
  MolecularHarmonicPotential  potentials.py:529-638
  MolecularGDMLPotential      potentials.py:641-744 + gdml_predictor.py:35-250
  RotatedMorsePotential       not in the reference: orthogonal change of coordinates around a Morse
                              potential, the dense-Hessian fixture of SURVEY.md section 8c-vi

There is no CPU path: harmonic_approximation() requires CUDA tensors.
"""
import ctypes
import logging

import numpy as np
import torch

from semiclassical_b200 import _native

__all__ = ['NonHarmonicPotential', 'MorsePotential', 'RotatedMorsePotential',
           'MolecularHarmonicPotential', 'MolecularGDMLPotential']

logger = logging.getLogger(__name__)


def _np64(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _NativePotential(object):
    """owns one sc_potential handle per CUDA device (created on first use on that device)"""
    _origin = 0.0
    _fused_step = True   # evaluated inside the fused step kernels (False: batched kernel + stage interface)

    def _create(self, out):
        raise NotImplementedError

    def _handle(self, device):
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError("semiclassical_b200 potentials run on CUDA devices only (no CPU fallback); got '%s'" % device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        handles = self.__dict__.setdefault('_handles', {})
        if idx not in handles:
            with torch.cuda.device(idx):
                h = ctypes.c_void_p()
                _native.check(self._create(ctypes.byref(h)))
                handles[idx] = h
        h = handles[idx]
        _native.check(_native.lib().sc_potential_set_origin(h, float(self._origin)))
        return h

    def __del__(self):
        try:
            for h in self.__dict__.get('_handles', {}).values():
                _native.lib().sc_potential_destroy(h)
        except Exception:
            pass

    def harmonic_approximation(self, r):
        """
        energies, gradients and Hessians for a batch of geometries

        r : real Tensor (dim, n) on a CUDA device  ->  vpot (n,), grad (dim, n), hess (dim, dim, n)
        """
        if not r.is_cuda:
            raise RuntimeError("semiclassical_b200: harmonic_approximation needs a CUDA tensor (no CPU fallback)")
        dim, n = r.shape
        assert dim == self.dimensions(), "position vectors have wrong dimensions"
        r = r.contiguous().to(torch.float64)
        vpot = torch.empty(n, dtype=torch.float64, device=r.device)
        grad = torch.empty((dim, n), dtype=torch.float64, device=r.device)
        hess = torch.empty((dim, dim, n), dtype=torch.float64, device=r.device)
        with torch.cuda.device(r.device):
            _native.check(_native.lib().sc_potential_eval(self._handle(r.device), n, r.data_ptr(), vpot.data_ptr(),
                                                          grad.data_ptr(), hess.data_ptr(), _stream_ptr(r.device)))
        return vpot, grad, hess

    def derivative_coupling_2nd(self, r):
        return torch.zeros_like(r)


class NonHarmonicPotential(_NativePotential):
    """eps*Morse + (1-eps)*harmonic, eqn. (7) of the Herman-Kluk paper (potentials.py:25-205)"""
    def __init__(self, eps=torch.tensor([0.975], dtype=torch.float64),
                 b=torch.tensor([(12.0)**(-0.5)], dtype=torch.float64)):
        self.eps = eps
        self.b = b

    def dimensions(self):
        return self.eps.size()[0]

    def masses(self):
        return torch.ones(self.dimensions(), dtype=torch.float64)

    def _create(self, out):
        eps, b = _np64(self.eps), _np64(self.b)
        return _native.lib().sc_potential_create_nonharmonic(out, len(eps), _p(eps), _p(b))

    def derivative_coupling_1st(self, r):
        return torch.ones_like(r)


class MorsePotential(_NativePotential):
    """V = sum_k D_k (1 - exp(-a_k r_k))^2 with a = sqrt(2 omega chi), D = omega/(4 chi) (potentials.py:208-397)"""
    def __init__(self, omega, chi, nac):
        self.omega = omega
        self.nac = nac
        if (chi == 0.0).all():
            logger.info("Potential is harmonic.")
        else:
            # harmonic modes get a tiny anharmonicity, in place like the reference (potentials.py:250)
            chi[chi == 0.0] += 1.0e-4
        self.chi = chi
        self.a = torch.sqrt(2 * omega * chi)
        self.D = 0.25 * omega / chi

    def dimensions(self):
        return self.a.size()[0]

    def masses(self):
        return torch.ones(self.dimensions(), dtype=torch.float64)

    def _arrays(self):
        omega, nac = _np64(self.omega), _np64(self.nac)
        allh = bool((self.chi == 0.0).all())
        a = np.zeros_like(omega) if allh else _np64(self.a)
        D = np.zeros_like(omega) if allh else _np64(self.D)
        return omega, a, D, int(allh), nac

    def _create(self, out):
        omega, a, D, allh, nac = self._arrays()
        return _native.lib().sc_potential_create_morse(out, len(omega), _p(omega), _p(a), _p(D), allh, _p(nac))

    def derivative_coupling_1st(self, r):
        return self.nac.to(r.device).unsqueeze(1).expand_as(r)


class RotatedMorsePotential(MorsePotential):
    """Morse potential in rotated coordinates x = Q r: V'(x) = V(Q^T x); dense Hessian Q h Q^T, tau1' = Q tau1"""
    def __init__(self, omega, chi, nac, Q):
        super().__init__(omega, chi, nac)
        self.Q = Q.to(torch.float64)
        self.nac_rot = self.Q @ self.nac.to(torch.float64)

    def _create(self, out):
        omega, a, D, allh, _ = self._arrays()
        nac, Q = _np64(self.nac_rot), _np64(self.Q)
        return _native.lib().sc_potential_create_rotated_morse(out, len(omega), _p(omega), _p(a), _p(D), allh, _p(nac), _p(Q))

    def derivative_coupling_1st(self, r):
        return self.nac_rot.to(r.device).unsqueeze(1).expand_as(r)


class _MolecularPotentialBase(_NativePotential):
    _masses, _dim = None, None

    def dimensions(self):
        return self._dim

    def masses(self):
        return self._masses

    def total_energy(self):
        return self._origin

    def derivative_coupling_1st(self, r):
        return self.nac0.to(r.device).unsqueeze(1).expand_as(r)

    def minimize(self, r_guess, maxiter=200, rtol=1.0e-5, gtol=1.0e-7, device=None):
        """
        Newton iteration with Armijo backtracking to the nearest minimum; shifts the energy origin there
        (same stopping rules as potentials.py:435-526; setup-time, one geometry)
        """
        device = torch.device(device if device is not None else ('cuda:%d' % torch.cuda.current_device()))
        self._origin = 0.0
        r = r_guess.to(device=device, dtype=torch.float64).unsqueeze(1)
        for it in range(maxiter):
            energy, grad, hess = self.harmonic_approximation(r)
            dr = torch.linalg.solve(hess[:, :, 0], -grad)
            slope = torch.sum(grad * dr)
            if slope > 0.0:
                dr = -grad
                slope = torch.sum(grad * dr)
            gnorm, dnorm = torch.norm(grad), torch.norm(dr)
            if gnorm < gtol or dnorm < rtol:
                break
            step = 1.0
            for _ in range(100):
                r_new = r + step * dr
                e_new = self.harmonic_approximation(r_new)[0]
                if e_new <= energy + 1.0e-4 * step * slope:
                    break
                step *= 0.3
            else:
                raise RuntimeError("Linesearch failed! Could not find a step length that satisfies the sufficient decrease condition.")
            r = r_new
        else:
            raise RuntimeError(f"Could not find minimum within {maxiter} iterations.")
        self._origin = self.harmonic_approximation(r)[0].item()
        logger.info(f"shift origin of energy axis to minimum energy = {self._origin} Hartree ")


class MolecularHarmonicPotential(_MolecularPotentialBase):
    """harmonic expansion around a reference geometry read from formatted checkpoint files (potentials.py:529-638)"""
    def __init__(self, freq_fchk, nac_fchk):
        pos0, energy0, grad0, hess0 = freq_fchk.harmonic_approximation()
        self._init_arrays(pos0, energy0, grad0, hess0, freq_fchk.masses(), nac_fchk.nonadiabatic_coupling())

    @classmethod
    def from_arrays(cls, pos0, energy0, grad0, hess0, masses, nac, origin=0.0):
        self = cls.__new__(cls)
        self._init_arrays(pos0, energy0, grad0, hess0, masses, nac)
        self._origin = float(origin)
        return self

    def _init_arrays(self, pos0, energy0, grad0, hess0, masses, nac):
        self.pos0 = torch.from_numpy(_np64(pos0))
        self.energy0 = torch.from_numpy(_np64(energy0).reshape(-1))
        self.grad0 = torch.from_numpy(_np64(grad0))
        self.hess0 = torch.from_numpy(_np64(hess0))
        self.nac0 = torch.from_numpy(_np64(nac))
        self._masses = torch.from_numpy(_np64(masses))
        self._dim = len(self._masses)

    def _create(self, out):
        pos0, grad0, hess0 = _np64(self.pos0), _np64(self.grad0), _np64(self.hess0)
        masses, nac = _np64(self._masses), _np64(self.nac0)
        return _native.lib().sc_potential_create_harmonic(out, self._dim, _p(pos0), float(self.energy0[0]), _p(grad0),
                                                          _p(hess0), _p(masses), _p(nac))


class MolecularGDMLPotential(_MolecularPotentialBase):
    """sGDML ground-state surface with analytic Hessians (potentials.py:641-744, gdml_predictor.py:35-250)"""

    @property
    def _fused_step(self):
        # 17 <= dim <= 64: the dense column pipeline drives k_gdml_eval itself (Hessians written straight into the stream
        # image of their RK4 stage); smaller / larger molecules go through the stage interface
        return 17 <= self._dim <= 64

    def __init__(self, model_pot, nac_fchk):
        model = dict(model_pot)
        assert np.array_equal(model['z'], nac_fchk.atomic_numbers()), \
            "GDML models for potential energy and NAC vector should be for the same molecule."
        self._init_model(model, nac_fchk.masses(), nac_fchk.nonadiabatic_coupling())

    @classmethod
    def from_arrays(cls, model, masses, nac, origin=0.0):
        self = cls.__new__(cls)
        self._init_model(dict(model), masses, nac)
        self._origin = float(origin)
        return self

    def _init_model(self, model, masses, nac):
        # expansion of the training set over the permutations (gdml_predictor.py:65-82)
        R_desc = _np64(model['R_desc'])
        alphas = _np64(np.array(model['R_d_desc_alpha']))
        n_desc = R_desc.shape[0]
        n_perms, n_atoms = np.asarray(model['perms']).shape
        perm_idxs = np.asarray(model['tril_perms_lin']).reshape(-1, n_perms).T
        self._xs_train = np.ascontiguousarray(np.tile(R_desc.T, (1, n_perms))[:, perm_idxs].reshape(-1, n_desc))
        self._jx_alphas = np.ascontiguousarray(np.tile(alphas, (1, n_perms))[:, perm_idxs].reshape(-1, n_desc))
        self._sig = float(int(model['sig']))
        self._c = float(model['c'])
        self._std = float(model.get('std', 1))
        self.n_atoms = int(n_atoms)
        self.nac0 = torch.from_numpy(_np64(nac))
        self._masses = torch.from_numpy(_np64(masses))
        self._dim = len(self._masses)
        assert self._dim == 3 * self.n_atoms

    def _create(self, out):
        masses, nac = _np64(self._masses), _np64(self.nac0)
        return _native.lib().sc_potential_create_gdml(out, self.n_atoms, self._xs_train.shape[0], self._xs_train.shape[1],
                                                      _p(self._xs_train), _p(self._jx_alphas), self._sig, self._c,
                                                      self._std, _p(masses), _p(nac))
