"""
CPU tests of the rates post-processing (semiclassical_b200/rates.py, row f4 of SURVEY section 8) against the reference's own
outputs (tests/golden/rates_as5.npz, written by oracle/make_golden.py from rates.rate_from_correlation) and the reference's
own unit test (tests/test_rates.py:15-46: the transform of the lineshape integrates to 1).
"""
import numpy as np

from helpers import load_golden, relerr
from semiclassical_b200 import rates, units


def test_rate_from_correlation_matches_reference():
    g = load_golden("rates_as5")
    e, r = rates.rate_from_correlation(g['times'], g['correlation'], rates.gaussian(float(g['sigma'])))
    assert np.array_equal(e, g['energies']) or relerr(e, g['energies']) < 1e-14
    assert relerr(r, g['rate_gaussian']) < 1e-12
    e, r = rates.rate_from_correlation(g['times'], g['correlation'], rates.lorentzian(float(g['gamma'])))
    assert relerr(e, g['energies_l']) < 1e-14 and relerr(r, g['rate_lorentzian']) < 1e-12


def test_lineshape_transform_is_normalised():
    hwhm_ev = 0.5
    sigma = hwhm_ev / np.sqrt(2.0 * np.log(2.0)) / units.hartree_to_ev
    times = np.linspace(0.0, 10.0, 500) / units.autime_to_fs
    w, G = rates.rate_from_correlation(times, np.ones_like(times), rates.gaussian(sigma))
    G = G / (1.0e15 / units.autime_to_fs)
    assert abs(np.sum(G * (w[1] - w[0])) - 1.0) < 1e-7
    # Voigt profile: product of the two transforms (convolution theorem)
    t = np.linspace(-3.0, 3.0, 7) / units.autime_to_fs
    v = rates.voigtian(sigma, 0.01)(t)
    assert relerr(v, 2.0 * np.pi * rates.gaussian(sigma)(t) * rates.lorentzian(0.01)(t)) < 1e-15


def test_calculate_rates_task(tmp_path):
    g = load_golden("rates_as5")
    f = tmp_path / "correlations.npz"
    np.savez(f, propagator="HK", times=g['times'], autocorrelation=g['correlation'], ic_correlation=g['correlation'], trajectories=1000)
    data = rates.calculate_rates({"task": "rates", "correlations": str(f), "rates": str(f), "hwhmG_ev": 0.01})
    keep = g['energies'] >= 0.0
    assert relerr(data['ic_rate'], (2.0 * np.pi * g['rate_gaussian'][keep]).real) < 1e-12
    out = dict(np.load(f))
    assert set(['energies', 'ic_rate', 'broadening', 'hwhmG', 'hwhmL']) <= set(out)
