"""
TEST INFRASTRUCTURE (never imported by the product path).

Runs the *unmodified* reference implementation (humeniuka/semiclassical, mounted read-only at
/root/reference) to produce golden vectors for tests/golden/.  Only usable in the build container:
the reference does not travel to the GPU box, the fixtures do.

Compatibility shim: torch.symeig / torch.solve were removed from torch; both are
re-created on top of torch.linalg before the reference is imported.  `ase` is absent; a minimal stub
(oracle/ase_stub) provides the few Atoms methods readers.py calls.
"""
import os
import sys
import logging
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("SEMICLASSICAL_REFERENCE", "/root/reference")


def _install_shims():
    torch.set_default_dtype(torch.float64)
    if not getattr(torch, "_sc_shim", False):
        def symeig(A, eigenvectors=False, upper=True):
            return torch.linalg.eigh(A, UPLO="U" if upper else "L")

        def solve(B, A):
            return torch.linalg.solve(A, B), None
        torch.symeig = symeig
        torch.solve = solve
        torch._sc_shim = True
    if "ase" not in sys.modules:
        _install_ase_stub()


def _install_ase_stub():
    """~30 line stand-in for ase.atoms.Atoms (only what readers.py / cli.py touch)."""
    class Atoms(object):
        def __init__(self, numbers=None):
            self.numbers = np.array(numbers)
            n = len(self.numbers)
            self.positions = np.zeros((n, 3))
            self.masses = np.ones(n)
            self.momenta = np.zeros((n, 3))

        def set_positions(self, pos): self.positions = np.array(pos, dtype=float).reshape(-1, 3)
        def get_positions(self): return self.positions.copy()
        def set_masses(self, m): self.masses = np.array(m, dtype=float)
        def get_masses(self): return self.masses.copy()
        def set_momenta(self, p): self.momenta = np.array(p, dtype=float).reshape(-1, 3)
        def get_center_of_mass(self): return self.masses @ self.positions / self.masses.sum()
        def translate(self, v): self.positions = self.positions + np.asarray(v)
        def copy(self):
            import copy
            return copy.deepcopy(self)

        def get_moments_of_inertia(self, vectors=False):
            com = self.get_center_of_mass()
            r = self.positions - com
            I = np.zeros((3, 3))
            for m, (x, y, z) in zip(self.masses, r):
                I += m * np.array([[y * y + z * z, -x * y, -x * z],
                                   [-x * y, x * x + z * z, -y * z],
                                   [-x * z, -y * z, x * x + y * y]])
            evals, evecs = np.linalg.eigh(I)
            return (evals, evecs.T) if vectors else evals

    ase = types.ModuleType("ase")
    ase.atoms = types.ModuleType("ase.atoms")
    ase.atoms.Atoms = Atoms
    ase.io = types.ModuleType("ase.io")
    ase.io.extxyz = types.ModuleType("ase.io.extxyz")
    ase.io.extxyz.write_extxyz = lambda *a, **k: None
    ase.Atoms = Atoms
    for name, mod in (("ase", ase), ("ase.atoms", ase.atoms), ("ase.io", ase.io), ("ase.io.extxyz", ase.io.extxyz)):
        sys.modules[name] = mod


def load_reference():
    """import the reference package read-only; returns (propagators, potentials, units, readers, gdml)"""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree %s not present (goldens can only be regenerated in the build container)"
                           % REFERENCE_ROOT)
    _install_shims()
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from semiclassical import propagators, potentials, units, readers, gdml_predictor
    logging.disable(logging.INFO)
    return propagators, potentials, units, readers, gdml_predictor


class RotatedPotential(object):
    """
    orthogonal change of coordinates x = Q r wrapped around a reference potential:
    V'(x) = V(Q^T x), grad' = Q grad, hess' = Q hess Q^T, tau1' = Q tau1.  Makes Hessian, Gamma and the
    monodromy blocks dense while leaving the correlation functions invariant.
    """
    def __init__(self, inner, Q):
        self.inner, self.Q = inner, Q

    def dimensions(self): return self.inner.dimensions()
    def masses(self): return self.inner.masses()

    def harmonic_approximation(self, x):
        r = self.Q.T @ x
        v, g, h = self.inner.harmonic_approximation(r)
        return v, self.Q @ g, torch.einsum('ai,ijn,bj->abn', self.Q, h, self.Q)

    def derivative_coupling_1st(self, x):
        return self.Q @ self.inner.derivative_coupling_1st(self.Q.T @ x)

    def derivative_coupling_2nd(self, x):
        return torch.zeros_like(x)


def run_reference(propagator, potential, dt, nt, energy0_es):
    """the cli.py:401-436 loop: read both correlations, then step"""
    auto = np.zeros(nt, dtype=complex)
    ic = np.zeros(nt, dtype=complex)
    for t in range(nt):
        auto[t] = propagator.autocorrelation(energy0_es=energy0_es)
        ic[t] = propagator.ic_correlation(potential, energy0_es=energy0_es)
        propagator.step(potential, dt)
    return auto, ic
