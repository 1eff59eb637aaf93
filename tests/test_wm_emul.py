"""
CPU check of the PRODUCT's Walton-Manolopoulos device routine (semiclassical_b200/csrc/sc_wm.cuh, a
__host__ __device__ template) compiled for the host with one thread per group (tests/emul/wm_emul.cu):
formulas, Gauss-Jordan inverses and scaled determinants against the oracle and the reference goldens.
The GPU parity tests (test_gpu_parity.py) run the same routine as a kernel.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import helpers
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "emul")
_dp = ctypes.POINTER(ctypes.c_double)


def _lib():
    so, src = os.path.join(EMUL, "libwm_emul.so"), os.path.join(EMUL, "wm_emul.cu")
    dep = os.path.join(os.path.dirname(HERE), "semiclassical_b200", "csrc", "sc_wm.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a",
                               "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def _emulate(g, consts, pot, y, c, signs):
    L = _lib()
    d, dr = consts.d, consts.dr
    k = consts._keep
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    U = f(k['U'].real)
    n1 = f(-pot.nac / pot.masses)
    cc = consts.c
    pref = np.sqrt(cc.detG0) * cc.detGt**0.25 * cc.detGi**0.25 / np.sqrt(cc.detGi0)
    n = y.shape[1]
    out4, det = np.zeros(4), np.zeros((n, 4))
    arrs = [f(k['Gamma_0']), f(k['Gamma_i']), f(k['Gamma_t']), f(k['iGi0']), f(k['iGamma_0']), U, f(k['q0']), f(k['p0']), n1]
    yy, zi, probi = f(y), f(g['zi'][:, :n]), f(g['probi'][:n])
    cv = np.ascontiguousarray(c, dtype=np.complex128)
    sg = f(signs)
    rc = L.wm_emul(ctypes.c_int(d), ctypes.c_int(dr), ctypes.c_int(n), *[a.ctypes.data_as(_dp) for a in arrs],
                   ctypes.c_double(cc.alpha), ctypes.c_double(cc.beta), ctypes.c_double(pref),
                   yy.ctypes.data_as(_dp), zi.ctypes.data_as(_dp), probi.ctypes.data_as(_dp), cv.ctypes.data_as(_dp),
                   sg.ctypes.data_as(_dp), out4.ctypes.data_as(_dp), det.ctypes.data_as(_dp))
    assert rc == 0
    return out4[0] + 1j * out4[1], out4[2] + 1j * out4[3], det[:, 0] + 1j * det[:, 1], det[:, 2] + 1j * det[:, 3]


@pytest.mark.parametrize("name,k", [("wm_1d", 9), ("wm_as5_chi002", 7), ("wm_as5_rot", 7), ("wm_methylium", 5)])
def test_wm_device_routine_on_host_matches_oracle(name, k):
    g = helpers.load_golden(name)
    pot, consts, wm = oracle.from_golden(g)
    assert wm
    n = min(len(g['probi']), 256)
    zi, probi = g['zi'][:, :n], g['probi'][:n]
    dt, e0 = float(g['dt']), float(g['energy0_es'])
    a = oracle.run(pot, consts, zi, probi, dt, k, e0, wm=True)          # state after k steps
    b = oracle.run(pot, consts, zi, probi, dt, k + 1, e0, wm=True)      # correlation sample k = at time k dt
    phase = np.exp(1j * k * dt * e0)
    ca, ki, detA, detM = _emulate(g, consts, pot, a['y'], a['c'], a['signs'])
    ref_a, ref_k = b['autocorrelation'][k] / phase * n, b['ic_correlation'][k] / phase * n
    assert abs(ca - ref_a) <= 1e-10 * abs(ref_a)
    assert abs(ki - ref_k) <= 1e-10 * abs(ref_k)


@pytest.mark.parametrize("name", ["wm_as5_chi002", "wm_as5_rot", "wm_methylium", "wm_1d"])
def test_wm_determinants_match_reference_golden(name):
    """det A / det M of the first trajectories at the final time of the fixture vs the unmodified reference"""
    g = helpers.load_golden(name)
    pot, consts, _ = oracle.from_golden(g)
    y = g['y_final']
    n = y.shape[1]
    signs = np.stack((g['signs_C'][:n], g['signs_detA'][:n].real, g['signs_detM'][:n].real)).real
    _, _, detA, detM = _emulate(g, consts, pot, y, g['c_final'][:n], signs)
    assert np.abs(detA - g['detA_final'][:n]).max() <= 1e-10 * np.abs(g['detA_final'][:n]).max()
    assert np.abs(detM - g['detM_final'][:n]).max() <= 1e-10 * np.abs(g['detM_final'][:n]).max()
