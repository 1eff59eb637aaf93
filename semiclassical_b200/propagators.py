"""
Herman-Kluk and Walton-Manolopoulos propagators with the interface of the reference's
semiclassical/propagators.py, backed by the fused sm_100a kernels (include/semiclassical_b200.h).

Host side (this file): argument checks, the O(d^3) setup algebra on the width matrices (eigendecompositions,
pseudo-inverses, overlap constants -- propagators.py:25-82, 125-179, 493-531, 1102-1130), sampling of the
initial ensemble (propagators.py:533-555) and the state machine  initial_conditions -> {autocorrelation,
ic_correlation, step}.  Device side: everything that touches a trajectory.

Beyond the reference interface:
  set_ensemble(zi, probi, ntraj_total)  install an externally sampled ensemble (parity runs, sharded ensembles)
  propagate(potential, dt, nsteps, energy0_es)  K fused steps per launch, returns both correlation functions
"""
import cmath
import ctypes
import logging

import numpy as np
import torch

from semiclassical_b200 import _native
from semiclassical_b200.units import hbar

__all__ = ['HermanKlukPropagator', 'WaltonManolopoulosPropagator']

ZERO = 1.0e-8   # threshold for treating singular values as 0 (propagators.py:16)

logger = logging.getLogger(__name__)


def _eigh(A):
    return torch.linalg.eigh(A.detach().to('cpu', torch.float64))


def _sym_sqrtm(A):
    """A^{1/2} (complex, all eigenvalues) and the pseudo-inverse A^{-1/2} (|e| > ZERO) of a symmetric real matrix"""
    e, V = _eigh(A)
    non_zero = abs(e) > ZERO
    e = e.type(torch.complex128)
    V = V.type(torch.complex128)
    sqA = torch.einsum('ij,j,kj->ik', V, torch.sqrt(e), V)
    sqA_pinv = torch.einsum('ij,j,kj->ik', V[:, non_zero], 1.0 / torch.sqrt(e[non_zero]), V[:, non_zero])
    return sqA, sqA_pinv


def _is_symmetric_non_negative(A, eps=1.0e-6):
    """A == A^T to relative accuracy eps and all eigenvalues >= -ZERO"""
    A = A.detach().to('cpu', torch.float64)
    relerr = torch.sum(abs(A - A.T)) / torch.sum(abs(A))
    if relerr > eps:
        return False
    e, _ = _eigh(A)
    return bool((e >= -ZERO).all())


def _pinv_sym(A):
    """pseudo-inverse, pseudo-determinant, rank of a symmetric matrix (eigenvalues with |e| <= ZERO dropped)"""
    e, V = _eigh(A)
    nz = abs(e) > ZERO
    inv = torch.einsum('ij,j,kj->ik', V[:, nz], 1.0 / e[nz], V[:, nz])
    return inv, torch.prod(e[nz]), int(torch.count_nonzero(nz))


class CoherentStatesOverlap(object):
    """overlap integrals <qi,pi,Gi|qj,pj,Gj> between batches of coherent states (propagators.py:124-240)"""
    def __init__(self, Gi, Gj):
        assert Gi.size() == Gj.size(), "width matrices Gi and Gj have to have the same shape"
        Gi = Gi.detach().to('cpu', torch.float64)
        Gj = Gj.detach().to('cpu', torch.float64)
        self.dim = Gi.size()[0]
        _, self.detGi, ranki = _pinv_sym(Gi)
        _, self.detGj, rankj = _pinv_sym(Gj)
        assert ranki == rankj, "Gi and Gj have to have the same rank and null space."
        self.Gij = Gi + Gj
        self.iGij, self.detGij, _ = _pinv_sym(self.Gij)
        self.Gi_iGij_Gj = Gi @ self.iGij @ Gj
        self.Gj_iGij = Gj @ self.iGij
        self.rank = ranki
        self.fac = torch.sqrt(2.0**self.rank * torch.sqrt(self.detGi) * torch.sqrt(self.detGj) / self.detGij)

    def __call__(self, qi, pi, qj, pj):
        """overlap matrix (ni, nj); diagnostic path (norm / wavefunction), plain torch on the tensors' device"""
        assert qi.size()[0] == pi.size()[0] == self.dim, "dimension of phase space points (qi, pi) is wrong"
        assert qj.size()[0] == pj.size()[0] == self.dim, "dimension of phase space points (qj, pj) is wrong"
        if qi.dim() == 1:
            qi, pi = qi.unsqueeze(1), pi.unsqueeze(1)
        if qj.dim() == 1:
            qj, pj = qj.unsqueeze(1), pj.unsqueeze(1)
        dev = qi.device
        A, B, C = (x.to(dev) for x in (self.Gi_iGij_Gj, self.iGij, self.Gj_iGij))
        dq = qj.unsqueeze(1) - qi.unsqueeze(2)      # (d, ni, nj)
        dp = pj.unsqueeze(1) - pi.unsqueeze(2)
        pjx = pj.unsqueeze(1).expand_as(dq)
        expo = (-0.5 * torch.einsum('aij,ab,bij->ij', dq, A, dq)
                - 0.5 / hbar**2 * torch.einsum('aij,ab,bij->ij', dp, B, dp)
                - 1j / hbar * torch.einsum('aij,aij->ij', pjx, dq)
                + 1j / hbar * torch.einsum('aij,ab,bij->ij', dq, C, dp))
        return self.fac.to(dev) * torch.exp(expo)


class CoherentStatesWavefunction(object):
    """phi(x) = sum_i v_i <x|q_i,p_i,G> on a spatial grid (propagators.py:241-290); plain torch on the tensors' device --
    the propagators evaluate the same sum with the tensor-core kernel behind `wavefunction(x)`"""
    def __init__(self, G):
        self.G = G
        _, self.detG, self.rank = _pinv_sym(G.detach().to('cpu', torch.float64))

    def __call__(self, q, p, v, x):
        d, nx = x.size()
        dim, ntraj = q.size()
        assert d == dim, "dimensions of spatial grid and coherent states differ"
        G = self.G.to(device=x.device, dtype=torch.float64)
        dx = x.unsqueeze(1) - q.unsqueeze(2)                      # (dim, ntraj, nx)
        fac = (self.detG / np.pi**self.rank)**0.25
        gaussians = fac * torch.exp(-0.5 * torch.einsum('inx,ij,jnx->nx', dx, G, dx)
                                    + 1j / hbar * torch.einsum('in,inx->nx', p, dx))
        return torch.sum(v.unsqueeze(1) * gaussians, 0)


def _np(x):
    return np.ascontiguousarray(x.detach().to('cpu', torch.float64).numpy())


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class HermanKlukPropagator(object):
    def __init__(self, Gamma_i, Gamma_t, device='cuda'):
        """
        semiclassical Herman-Kluk propagator (propagators.py:407-443)

        Gamma_i, Gamma_t : real symmetric positive semi-definite Tensors (dim,dim), widths of the frozen
                           Gaussians at t=0 and at later times
        device           : CUDA device all trajectory data lives on
        """
        assert _is_symmetric_non_negative(Gamma_i), "Gamma_i has to be symmetric and positive semi-definite."
        assert _is_symmetric_non_negative(Gamma_t), "Gamma_t has to be symmetric and positive semi-definite."
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError("semiclassical_b200 propagators run on CUDA devices only (no CPU fallback); got device='%s'" % device)
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = device
        _native.lib()   # fail now, not at the first step, if the CUDA library is missing
        self.Gamma_i = Gamma_i.to(device=device, dtype=torch.float64)
        self.Gamma_t = Gamma_t.to(device=device, dtype=torch.float64)
        self.sqGi, self.isqGi = _sym_sqrtm(Gamma_i)
        self.sqGt, self.isqGt = _sym_sqrtm(Gamma_t)
        self._engine = None
        self._energies = []
        self._corr_cache = None
        self._wm = 0
        self.alpha = self.beta = torch.tensor(1.0)

    # ------------------------------------------------------------------ setup
    def __del__(self):
        try:
            if self._engine is not None:
                _native.lib().sc_engine_destroy(self._engine)
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _setup_constants(self, q0, p0, Gamma_0):
        """everything that depends only on the width matrices and the wavepacket centre (host, fp64)"""
        G0 = Gamma_0.detach().to('cpu', torch.float64)
        Gi = self.Gamma_i.detach().to('cpu', torch.float64)
        Gt = self.Gamma_t.detach().to('cpu', torch.float64)
        d = G0.shape[0]
        # same wavepacket as last time (a new repetition of the same run): keep the engine and its device buffers
        key = (_np(q0).tobytes(), _np(p0).tobytes(), _np(G0).tobytes(), self._wm, float(self.alpha), float(self.beta))
        if self._engine is not None and getattr(self, '_const_key', None) == key:
            return self._const_dims
        # non-zero subspace of Gi + G0 and its pseudo-inverse (propagators.py:493-501)
        wp, Vp = _eigh(Gi + G0)
        nzp = wp > ZERO
        U = Vp[:, nzp]
        iGi0 = torch.einsum('ij,j,kj->ik', U, 1.0 / wp[nzp], U)
        self.U = U.type(torch.complex128).to(self.device)
        self.iGi0 = iGi0.to(self.device)
        dr = int(nzp.sum())
        # prefactor factors with the projection folded in; for positive semi-definite widths they are real
        Uc = U.type(torch.complex128)
        facs = [Uc.T @ self.sqGt, Uc.T @ self.isqGt, self.isqGi @ Uc, self.sqGi @ Uc]
        for f in facs:
            if float(abs(f.imag).max()) > 1.0e-12 * max(float(abs(f.real).max()), 1.0e-300):
                raise NotImplementedError("width matrices with negative eigenvalues inside the propagated subspace are not supported")
        L1, L2, R1, R2 = (np.ascontiguousarray(f.real.numpy()) for f in facs)
        self.csoi0 = CoherentStatesOverlap(Gi, G0)
        self.csot0 = CoherentStatesOverlap(Gt, G0)
        self.csott = CoherentStatesOverlap(Gt, Gt)
        keep = dict(L1=L1, L2=L2, R1=R1, R2=R2, U=_np(U), q0=_np(q0), p0=_np(p0),
                    oi0_A=_np(self.csoi0.Gi_iGij_Gj), oi0_B=_np(self.csoi0.iGij), oi0_C=_np(self.csoi0.Gj_iGij),
                    ot0_A=_np(self.csot0.Gi_iGij_Gj), ot0_B=_np(self.csot0.iGij), ot0_C=_np(self.csot0.Gj_iGij),
                    Gamma_0=_np(G0), Gamma_i=_np(Gi), Gamma_t=_np(Gt), iGi0=_np(iGi0))
        cfg = _native.EngineConfig()
        cfg.d, cfg.dr, cfg.wm = d, dr, self._wm
        cfg.oi0_fac, cfg.ot0_fac = float(self.csoi0.fac), float(self.csot0.fac)
        cfg.alpha, cfg.beta = float(self.alpha), float(self.beta)
        if self._wm:
            # pi-absorbed pseudo-determinants and pseudo-inverse of Gamma_0 (propagators.py:1117-1130)
            def pdet(G, scale):
                e, _ = _eigh(G)
                return float(torch.prod(e[abs(e) > ZERO] / scale))
            self.detG0, self.detGi, self.detGt = pdet(G0, np.pi), pdet(Gi, np.pi), pdet(Gt, np.pi)
            self.detGi0 = pdet(G0 + Gi, 2 * np.pi)
            e0, V0 = _eigh(G0)
            nz0 = e0 > ZERO
            self.iGamma_0 = torch.einsum('ij,j,kj->ik', V0[:, nz0], 1.0 / e0[nz0], V0[:, nz0])
            keep['iGamma_0'] = _np(self.iGamma_0)
            cfg.detG0, cfg.detGi, cfg.detGt, cfg.detGi0 = self.detG0, self.detGi, self.detGt, self.detGi0
        for name, arr in keep.items():
            setattr(cfg, name, _ptr(arr))
        if self._engine is not None:
            _native.lib().sc_engine_destroy(self._engine)
            self._engine = None
        eng = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_create(ctypes.byref(eng), ctypes.byref(cfg)))
        self._engine = eng
        self._iLz_detLz = None
        self._const_key, self._const_dims = key, (d, dr)
        return d, dr

    def initial_conditions(self, q0, p0, Gamma_0, ntraj=5000, ntraj_total=None, index0=0, seed=None):
        """
        sample initial positions and momenta from P(qi,pi) ~ |<qi,pi,Gamma_i|q0,p0,Gamma_0>|^2
        (propagators.py:445-631) and install them on the device

        ntraj_total : size of the global ensemble when this propagator holds one shard of it
        """
        assert Gamma_0.size() == self.Gamma_i.size(), "Width parameter matrix Gamma_0 has wrong dimensions."
        assert _is_symmetric_non_negative(Gamma_0), "Gamma_0 has to be symmetric and positive semi-definite."
        zi, probi = self.sample_ensemble(q0, p0, Gamma_0, ntraj, index0=index0, seed=seed)
        self._install(zi, probi, ntraj_total)

    def sample_ensemble(self, q0, p0, Gamma_0, ntraj, index0=0, seed=None):
        """
        draws ntraj phase-space points from P(qi,pi) ~ |<qi,pi,Gamma_i|q0,p0,Gamma_0>|^2 (propagators.py:493-555) WITHOUT
        installing them: returns zi (2 dim, ntraj), probi (ntraj,) on the device.  Sampling runs in the engine's own kernel
        (Philox4x32-10 counter-based generator + Box-Muller, k_sample_ensemble): the ensemble is a pure function of
        (seed, index0 + i).  seed=None draws the 64-bit seed from torch's default CPU generator, so torch.manual_seed(s)
        makes runs reproducible as in the reference; ranks that seed identically and pass the first global index of their
        shard as index0 draw slices of ONE global ensemble.
        """
        self._prepare(q0, p0, Gamma_0)
        d = self.dim
        G0 = Gamma_0.detach().to('cpu', torch.float64)
        Gi = self.Gamma_i.detach().to('cpu', torch.float64)
        wp, Vp = _eigh(Gi + G0)
        nzp = wp > ZERO
        iLp = torch.einsum('i,ji->ij', torch.sqrt(wp[nzp] / 2), Vp[:, nzp])
        wq, Vq = _eigh(Gi @ self.iGi0.cpu() @ G0)
        nzq = wq > ZERO
        iLq = torch.einsum('i,ji->ij', 1.0 / torch.sqrt(2 * wq[nzq]), Vq[:, nzq])
        assert int(nzp.sum()) == int(nzq.sum()), \
            "number of non-zero modes for sampling of positions and momenta have to be the same"
        nnz = int(nzp.sum())
        assert nnz == self.rank
        detLz = torch.prod(2 * torch.sqrt(wq[nzq] / wp[nzp])).item()
        if seed is None:
            seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())
        zi = torch.empty((2 * d, ntraj), dtype=torch.float64, device=self.device)
        probi = torch.empty(ntraj, dtype=torch.float64, device=self.device)
        a_q, a_p = _np(iLq), _np(iLp)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_sample_ensemble(self._engine, int(ntraj), int(index0), int(seed), _ptr(a_q), _ptr(a_p),
                                                                  float(detLz), zi.data_ptr(), probi.data_ptr(), self._stream()))
        self._iLz_detLz = (iLq, iLp, detLz)
        logger.info("== Initial Conditions ==")
        logger.info(f"number of dimensions   :  {d}")
        logger.info(f"zero dimensions        :  {d - nnz}")
        logger.info(f"number of trajectories :  {ntraj}")
        return zi, probi

    def set_ensemble(self, q0, p0, Gamma_0, zi, probi, ntraj_total=None):
        """install an externally sampled ensemble: zi (2 dim, n), probi (n,) (ensemble injection)"""
        assert Gamma_0.size() == self.Gamma_i.size(), "Width parameter matrix Gamma_0 has wrong dimensions."
        assert _is_symmetric_non_negative(Gamma_0), "Gamma_0 has to be symmetric and positive semi-definite."
        self._prepare(q0, p0, Gamma_0)
        assert zi.shape[0] == 2 * self.dim and zi.shape[1] == probi.shape[0], "ensemble has wrong shape"
        self._install(zi, probi, ntraj_total)

    def _prepare(self, q0, p0, Gamma_0):
        self.q0 = q0.to(device=self.device, dtype=torch.float64)
        self.p0 = p0.to(device=self.device, dtype=torch.float64)
        self.Gamma_0 = Gamma_0.to(device=self.device, dtype=torch.float64)
        self.dim, self.rank = self._setup_constants(q0, p0, Gamma_0)

    def _install(self, zi, probi, ntraj_total):
        self.zi = zi.to(device=self.device, dtype=torch.float64).contiguous()
        self.probi = probi.to(device=self.device, dtype=torch.float64).contiguous()
        self.ntraj = int(self.probi.shape[0])
        self.ntraj_total = int(ntraj_total) if ntraj_total else self.ntraj
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_set_ensemble(self._engine, self.ntraj, self.ntraj_total,
                                                               self.zi.data_ptr(), self.probi.data_ptr(), self._stream()))
        self.t = 0.0
        self._energies = []
        self._corr_cache = None

    # ------------------------------------------------------------------ propagation
    def _native_potential(self, potential):
        handle = getattr(potential, '_handle', None)
        return handle(self.device) if handle is not None else None

    def _fused_potential(self, potential):
        """handle of a potential the fused step kernels evaluate in-kernel, else None (stage interface)"""
        if not getattr(potential, '_fused_step', True):
            return None
        return self._native_potential(potential)

    def _check_energy(self, energies, change_tol=1.0e-2):
        """<T+V> (4th RK4 stage) must not change by more than change_tol between steps (propagators.py:385-398)"""
        for en in energies:
            self._energies.append(float(en))
            if len(self._energies) > 1:
                change = abs(self._energies[1] - self._energies[0])
                if change > change_tol:
                    logger.error("  energy conservation violated")
                    raise RuntimeError(f"average energy of classical trajectories is not conserved, change= {change} Hartree")
                self._energies.pop(0)

    def step(self, potential, dt):
        """propagates the ensemble for one time step (t -> t+dt) under the influence of `potential`"""
        assert self.dim == potential.dimensions(), "potential has wrong dimensions"
        h = float(dt)
        handle = self._fused_potential(potential)
        if handle is None:
            self._step_generic(potential, h)
        else:
            out = np.zeros(5)
            with torch.cuda.device(self.device):
                _native.check(_native.lib().sc_engine_step(self._engine, handle, h, 1, out.ctypes.data, self._stream()))
            self._check_energy(out[4:5])
            self._corr_cache = (potential, out[0] + 1j * out[1], out[2] + 1j * out[3])
        self.t += dt

    def propagate(self, potential, dt, nsteps, energy0_es=0.0, group=None):
        """
        nsteps fused time steps in one launch.  Returns (autocorrelation, ic_correlation), complex arrays of
        length nsteps holding the values at the nsteps NEW times t+dt ... t+nsteps*dt, including the
        e^{i t E0 / hbar} phase -- the same numbers nsteps x {step; autocorrelation; ic_correlation} produce.

        group : torch.distributed process group (or True for the default group) when this propagator holds one
                shard of a global ensemble: the per-step sums are combined by ONE all-reduce on the device
                (semiclassical_b200.distributed) before they are copied to the host.
        """
        assert self.dim == potential.dimensions(), "potential has wrong dimensions"
        handle = self._fused_potential(potential)
        if handle is None:
            # potentials evaluated outside the fused kernels (Python objects, sGDML): step by step
            assert group is None, "sharded propagation needs a potential the fused kernels evaluate"
            auto, ic = np.zeros(nsteps, complex), np.zeros(nsteps, complex)
            for k in range(nsteps):
                self.step(potential, dt)
                auto[k] = self.autocorrelation(energy0_es)
                ic[k] = self.ic_correlation(potential, energy0_es)
            return auto, ic
        from semiclassical_b200 import distributed
        h = float(dt)
        rows = torch.empty((nsteps, 5), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_step_dev(self._engine, handle, h, nsteps, rows.data_ptr(), self._stream()))
            local_energy = rows[:, 4].clone() if group is not None else None
            if group is not None:
                distributed.allreduce_rows(rows, self.ntraj, self.ntraj_total, None if group is True else group)
            out = rows.cpu().numpy()
        self._check_energy(out[:, 4] if local_energy is None else local_energy.cpu().numpy())
        times = np.zeros(nsteps)
        t = self.t
        for k in range(nsteps):
            t = t + dt
            times[k] = float(t)
        self.t = t
        auto, ic = distributed.rows_to_correlations(out, times, energy0_es, hbar)
        if group is None:
            self._corr_cache = (potential, out[-1, 0] + 1j * out[-1, 1], out[-1, 2] + 1j * out[-1, 3])
        else:
            self._corr_cache = None   # the cached values must stay local sums
        return auto, ic

    def _correlations(self, potential):
        """(C_auto, k_ic) sums at the current time without the dynamical phase"""
        if self._corr_cache is not None and (potential is None or self._corr_cache[0] is potential):
            return self._corr_cache[1], self._corr_cache[2]
        out = np.zeros(4)
        with torch.cuda.device(self.device):
            if potential is None:
                n1 = np.zeros(self.dim)
                _native.check(_native.lib().sc_engine_correlations_n1(self._engine, n1.ctypes.data, out.ctypes.data, self._stream()))
            else:
                handle = self._native_potential(potential)
                if handle is not None:
                    _native.check(_native.lib().sc_engine_correlations(self._engine, handle, out.ctypes.data, self._stream()))
                else:
                    n1 = self._constant_n1(potential)
                    if n1 is not None:
                        _native.check(_native.lib().sc_engine_correlations_n1(self._engine, n1.ctypes.data, out.ctypes.data, self._stream()))
                    else:
                        self._correlations_general(potential, out)
        res = (out[0] + 1j * out[1], out[2] + 1j * out[3])
        if potential is not None:
            self._corr_cache = (potential, res[0], res[1])
        return res

    def _constant_n1(self, potential):
        """n1 = -hbar^2 tau1/m for potentials whose NAC vector is constant and whose second-order coupling vanishes (all shipped
        ones, Condon approximation); None if the couplings depend on the position (-> _correlations_general)"""
        q, _ = self.current_positions_and_momenta()
        probe = q[:, :min(3, self.ntraj)]
        tau1 = potential.derivative_coupling_1st(probe)
        tau2 = potential.derivative_coupling_2nd(probe)
        varies = tau1.shape[1] > 1 and not all(torch.equal(tau1[:, 0], tau1[:, k]) for k in range(1, tau1.shape[1]))
        if not varies and self.ntraj > 1:
            # a second probe at the initial positions: constant means constant everywhere the ensemble has been
            tau1i = potential.derivative_coupling_1st(self.zi[:self.dim, :1].contiguous())
            varies = not torch.equal(tau1i[:, 0], tau1[:, 0])
        if varies or float(abs(tau2).max()) != 0.0:
            return None
        masses = potential.masses().to(self.device)
        return np.ascontiguousarray((-hbar**2 * tau1[:, 0] / masses).detach().cpu().numpy())

    def _correlations_general(self, potential, out):
        """position-dependent couplings (propagators.py:868-909 in full generality): the potential object evaluates tau1, tau2 at
        the initial and at the current positions, the engine's kernel does the rest (k_corr_general)"""
        if self._wm:
            raise NotImplementedError("position-dependent non-adiabatic couplings: Herman-Kluk propagator only")
        d = self.dim
        Q, _ = self.current_positions_and_momenta()
        Q = Q.contiguous()
        q = self.zi[:d].contiguous()
        im = (1.0 / potential.masses().to(device=self.device, dtype=torch.float64)).unsqueeze(1)
        n1Q = (-hbar**2 * im * potential.derivative_coupling_1st(Q)).to(torch.float64).contiguous()
        n1q = (-hbar**2 * im * potential.derivative_coupling_1st(q)).to(torch.float64).contiguous()
        n2Q = (-0.5 * hbar**2 * (im * potential.derivative_coupling_2nd(Q)).sum(dim=0)).to(torch.float64).contiguous()
        n2q = (-0.5 * hbar**2 * (im * potential.derivative_coupling_2nd(q)).sum(dim=0)).to(torch.float64).contiguous()
        _native.check(_native.lib().sc_engine_correlations_general(self._engine, n1Q.data_ptr(), n1q.data_ptr(), n2Q.data_ptr(),
                                                                   n2q.data_ptr(), out.ctypes.data, self._stream()))

    def autocorrelation(self, energy0_es=0.0):
        """e^{i t E0/hbar} <phi(0)|phi(t)> at the current time step (propagators.py:809-843)"""
        cauto, _ = self._correlations(None if self._corr_cache is None else self._corr_cache[0])
        return complex(cauto * cmath.exp(1j / hbar * float(self.t) * energy0_es))

    def ic_correlation(self, potential, energy0_es=0.0):
        """correlation function for the internal-conversion rate at the current time step (propagators.py:845-911)"""
        _, kic = self._correlations(potential)
        return complex(kic * cmath.exp(1j / hbar * float(self.t) * energy0_es))

    # ------------------------------------------------------------------ generic potentials
    def _step_generic(self, potential, h):
        """RK4 step with a user-supplied Python potential: the potential is evaluated by its own
        harmonic_approximation(), the monodromy/prefactor work stays in the kernels"""
        L = _native.lib()
        d, n = self.dim, self.ntraj
        masses = potential.masses().to(device=self.device, dtype=torch.float64).contiguous()
        q = torch.empty((d, n), dtype=torch.float64, device=self.device)
        esum = torch.zeros(1, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            for stage in range(1, 5):
                _native.check(L.sc_engine_stage_positions(self._engine, stage, h, q.data_ptr(), self._stream()))
                vpot, grad, hess = potential.harmonic_approximation(q)
                vpot, grad, hess = (x.to(torch.float64).contiguous() for x in (vpot, grad, hess))
                _native.check(L.sc_engine_stage_apply(self._engine, stage, h, masses.data_ptr(), vpot.data_ptr(),
                                                      grad.data_ptr(), hess.data_ptr(), esum.data_ptr(), self._stream()))
            _native.check(L.sc_engine_stage_finish(self._engine, h, self._stream()))
        self._check_energy([esum.item() / n])
        self._corr_cache = None

    # ------------------------------------------------------------------ data access (propagators.py:914-948)
    @property
    def y(self):
        """solution vector in the reference's layout (2 dim + 4 dim^2 + 1, ntraj): q, p, Mqq, Mqp, Mpq, Mpp, S"""
        d = self.dim
        y = torch.empty((2 * d + 4 * d * d + 1, self.ntraj), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_get_state(self._engine, y.data_ptr(), self._stream()))
        return y

    @y.setter
    def y(self, value):
        d = self.dim
        value = value.to(device=self.device, dtype=torch.float64).contiguous()
        assert value.shape == (2 * d + 4 * d * d + 1, self.ntraj), "solution vector has wrong shape"
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_set_state(self._engine, value.data_ptr(), self._stream()))
        self._corr_cache = None

    def _prefactor_arrays(self):
        n = self.ntraj
        c = torch.empty(n, dtype=torch.complex128, device=self.device)
        c2 = torch.empty(n, dtype=torch.complex128, device=self.device)
        signs = torch.empty((3, n), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_get_prefactor(self._engine, c.data_ptr(), c2.data_ptr(), signs.data_ptr(),
                                                                self._stream()))
        return c, c2, signs

    @property
    def c(self):
        """sqrt(det) on the principal branch, without alignment of signs"""
        return self._prefactor_arrays()[0]

    @property
    def sign_trackers(self):
        c, c2, signs = self._prefactor_arrays()
        return {"prefactorC": {"signs": signs[0].type(torch.complex128), "previous": c2}}

    def initial_positions_and_momenta(self):
        d = self.dim
        return torch.split(self.zi, [d, d])

    def current_positions_and_momenta(self):
        d = self.dim
        q, p = torch.split(self.y, [d, d, d**2, d**2, d**2, d**2, 1])[:2]
        return q, p

    def classical_action(self):
        return self.y[-1]

    def monodromy_matrices(self):
        d = self.dim
        _, _, Mqq, Mqp, Mpq, Mpp, _ = torch.split(self.y, [d, d, d**2, d**2, d**2, d**2, 1])
        return Mqq.view(d, d, -1), Mqp.view(d, d, -1), Mpq.view(d, d, -1), Mpp.view(d, d, -1)

    def semiclassical_prefactor(self):
        """prefactor C(t) with the branch signs applied"""
        c, _, signs = self._prefactor_arrays()
        return signs[0] * c

    def autocorrelation_qp(self):
        """contribution of each trajectory to the autocorrelation function, <phi(0)|qt,pt> C(t) e^{iS} <qi,pi|phi(0)>
        (propagators.py:784-807).  Diagnostic accessor: torch operations on the exported state; the correlation functions
        themselves are accumulated inside the step kernels."""
        qi, pi = self.initial_positions_and_momenta()
        qt, pt = self.current_positions_and_momenta()
        q0, p0 = self.q0.to(self.device), self.p0.to(self.device)
        vi = self.csoi0(qi, pi, q0, p0).squeeze()
        vt = self.csot0(qt, pt, q0, p0).squeeze()
        return vt.conj() * vi * self.semiclassical_prefactor() * torch.exp(1j / hbar * self.classical_action())

    # ------------------------------------------------------------------ wavefunction diagnostics (propagators.py:657-782)
    def coefficients(self):
        """
        expansion coefficients of the Herman-Kluk wavefunction in the basis of the coherent state trajectories
        (propagators.py:657-686):  v_i = C_i e^{i S_i} / (2 pi)^d <q_i,p_i|phi(0)> / (ntraj P_i), complex Tensor (ntraj,)
        """
        v = torch.empty(self.ntraj, dtype=torch.complex128, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_coefficients(self._engine, v.data_ptr(), self._stream()))
        return v

    def wavefunction(self, x):
        """
        frozen Gaussian approximation of the wavefunction psi(x,t) on a spatial grid (propagators.py:688-732)

        x : real Tensor (dim,nx)   ->   complex numpy.ndarray (nx,)
        """
        d, nx = x.shape
        assert d == self.dim, "spatial grid has wrong dimensions"
        Gt = self.Gamma_t.detach().to('cpu', torch.float64)
        _, detG, rank = _pinv_sym(Gt)
        fac = float((detG / np.pi**rank)**0.25)
        xg = x.to(device=self.device, dtype=torch.float64).contiguous()
        phi = torch.empty(nx, dtype=torch.complex128, device=self.device)
        G = _np(Gt)
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_wavefunction(self._engine, _ptr(G), fac, nx, xg.data_ptr(), phi.data_ptr(),
                                                               self._stream()))
        return phi.detach().cpu().numpy()

    def norm(self, group=None):
        """
        norm |psi| = sqrt(sum_ij v_i^* v_j <g_i|g_j>) of the frozen Gaussian wavefunction (propagators.py:734-782);
        all pairs of trajectories, O(ntraj^2), as two FP64 tensor-core contractions + exp/sincos per pair.

        group : torch.distributed process group (or True for the default group) when this propagator holds one shard of a
                global ensemble: the ket vectors of all shards are all-gathered, every rank sums its
                (n_local x n_total) block of pairs, and the blocks are all-reduced -- the norm of the GLOBAL wavefunction
                on every rank (cli.py:424-429 computes it on the one process that holds everything)
        """
        cs = CoherentStatesOverlap(self.Gamma_t, self.Gamma_t)
        A, B, C = _np(cs.Gi_iGij_Gj), _np(cs.iGij), _np(cs.Gj_iGij)
        out = np.zeros(2)
        L = _native.lib()
        if group is None:
            with torch.cuda.device(self.device):
                _native.check(L.sc_engine_norm(self._engine, _ptr(A), _ptr(B), _ptr(C), float(cs.fac), _ptr(out), self._stream()))
            return float(np.sqrt(out[0]))
        import torch.distributed as dist
        pg = None if group is True else group
        world, rank = dist.get_world_size(pg), dist.get_rank(pg)
        counts = torch.zeros(world, dtype=torch.int64, device=self.device)
        counts[rank] = self.ntraj
        dist.all_reduce(counts, group=pg)
        counts = [int(c) for c in counts.cpu()]
        n_pad = max(counts)
        pack = self.norm_pack(n_pad)
        packs = torch.empty(world * pack.numel(), dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(packs, pack, group=pg)
        tot = 0.0j
        for r in range(world):
            tot += self.norm_block(packs[r * pack.numel():(r + 1) * pack.numel()], counts[r], n_pad, pack)
        tot = torch.tensor([tot.real, tot.imag], dtype=torch.float64, device=self.device)
        dist.all_reduce(tot, group=pg)
        return float(np.sqrt(float(tot[0])))

    def norm_pack(self, n_pad):
        """ket vectors of this shard for the sharded norm (include/semiclassical_b200.h: sc_engine_norm_pack); 1-D fp64 tensor"""
        cs = CoherentStatesOverlap(self.Gamma_t, self.Gamma_t)
        A, B, C = _np(cs.Gi_iGij_Gj), _np(cs.iGij), _np(cs.Gj_iGij)
        self._norm_fac = float(cs.fac)
        size = ctypes.c_longlong()
        L = _native.lib()
        _native.check(L.sc_engine_norm_pack_size(self._engine, n_pad, ctypes.byref(size)))
        pack = torch.zeros(size.value, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(L.sc_engine_norm_pack(self._engine, _ptr(A), _ptr(B), _ptr(C), n_pad, pack.data_ptr(), self._stream()))
        return pack

    def norm_block(self, other_pack, n_ket, n_pad, own_pack):
        """sum_{i in this shard} conj(v_i) sum_{j in other_pack} <g_i|g_j> v_j  (complex); norm_pack() must have been called"""
        kp = (2 * self.dim + 3) & ~3
        out = np.zeros(2)
        other_pack = other_pack.contiguous()
        with torch.cuda.device(self.device):
            _native.check(_native.lib().sc_engine_norm_block(self._engine, int(n_ket), int(n_pad), other_pack.data_ptr(),
                                                             own_pack.data_ptr() + 8 * n_pad * (2 * kp + 2), self._norm_fac,
                                                             _ptr(out), self._stream()))
        return complex(out[0], out[1])

    def set_option(self, name, value):
        """engine run-time options (include/semiclassical_b200.h: sc_engine_set_option), e.g. 'dense_engine'"""
        _native.check(_native.lib().sc_engine_set_option(self._engine, name.encode(), int(value)))

    def launch_count(self):
        return int(_native.lib().sc_engine_launch_count(self._engine))

    def kernel_name(self):
        return _native.lib().sc_engine_kernel_name(self._engine).decode()


class WaltonManolopoulosPropagator(HermanKlukPropagator):
    def __init__(self, Gamma_i, Gamma_t, alpha, beta, device='cuda'):
        """
        Walton-Manolopoulos propagator (propagators.py:1077-1100): the HK propagator integrated over a
        phase-space cell of width ~ (2 alpha)^(-dim/2) in position and (2 beta)^(-dim/2) in momentum
        """
        super().__init__(Gamma_i, Gamma_t, device=device)
        self.alpha = torch.tensor(alpha)
        self.beta = torch.tensor(beta)
        self._wm = 1

    def autocorrelation_qp(self):
        # the per-trajectory WM contributions (eqn 85, propagators.py:1577-1632) exist only inside k_wm / k_wm_fused
        raise NotImplementedError("per-trajectory contributions are not exported for the Walton-Manolopoulos propagator; "
                                  "autocorrelation() and ic_correlation() return the sums")

    @property
    def sign_trackers(self):
        c, c2, signs = self._prefactor_arrays()
        return {"prefactorC": {"signs": signs[0].type(torch.complex128), "previous": c2},
                "detA": {"signs": signs[1].type(torch.complex128)},
                "detM": {"signs": signs[2].type(torch.complex128)}}
