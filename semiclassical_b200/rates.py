"""
Rates task of `semi` (row f4 of SURVEY.md section 8): k_IC(E) as the Fourier transform of the internal-conversion
correlation function that the dynamics task accumulated (reference: rates.py:20-82 `rate_from_correlation`,
broadening.py:25-146 lineshape factories, cli.py:519-570 `calculate_rates`).  O(nt log nt) host post-processing on a few
thousand points: plain numpy, not a GPU path.

    k(E) = 1/(2 pi hbar)  Int dt  e^{i E t / hbar}  f(t)  k~(t)

k~ is known on [0, t_max]; negative times follow from k~(-t) = conj(k~(t)).  The integrand is multiplied by the lineshape
f(t) (Fourier transform of the broadening function) and by a cos^2 switching function that takes it to zero at |t| = t_max.
"""
import logging

import numpy as np

from semiclassical_b200 import units

__all__ = ['gaussian', 'lorentzian', 'voigtian', 'rate_from_correlation', 'calculate_rates']

logger = logging.getLogger(__name__)


def gaussian(sigma):
    """g(t) = exp(-sigma^2 t^2 / 2) / (2 pi): transform of a normalised Gaussian of standard deviation sigma (broadening.py:25-65)"""
    return lambda t: np.exp(-0.5 * (sigma * np.asarray(t, dtype=float))**2) / (2.0 * np.pi)


def lorentzian(gamma):
    """l(t) = exp(-gamma |t|) / (2 pi) for t != 0 and 0 at t = 0 exactly as the reference evaluates it (broadening.py:67-113:
    the two one-sided exponentials are only assigned for t > 0 and t < 0)"""
    def lineshape(t):
        t = np.asarray(t, dtype=float)
        return np.where(t != 0.0, np.exp(-gamma * np.abs(t)), 0.0) / (2.0 * np.pi)
    return lineshape


def voigtian(sigma, gamma):
    """product of the Gaussian and Lorentzian transforms (convolution theorem), 2 pi v(t) = (2 pi g(t)) (2 pi l(t)).
    (The reference's voigtian, broadening.py:115-146, calls its factories with the wrong arity and raises.)"""
    g, l = gaussian(sigma), lorentzian(gamma)
    return lambda t: 2.0 * np.pi * g(t) * l(t)


def rate_from_correlation(times, correlation, lineshape):
    """
    times (nt,) equidistant from 0 to t_max, correlation (nt,) complex, lineshape callable  ->  energies (2 nt - 1,) in
    Hartree (ascending), rate (2 nt - 1,) complex in s^-1.  Same discretisation as rates.py:20-82: 2 nt - 1 samples on
    [-t_max, t_max], inverse DFT scaled by the length of the interval.
    """
    times = np.asarray(times, dtype=float)
    correlation = np.asarray(correlation, dtype=complex)
    assert times.min() == 0.0, "time grid `times` should start at 0.0"
    assert times.shape == correlation.shape, "arrays `times` and `correlation` should have the same length"
    nt = times.shape[0]
    t_max = times.max()
    m = 2 * nt - 1
    t_full = np.linspace(-t_max, t_max, m)
    k_full = np.concatenate((np.conj(correlation[:0:-1]), correlation))           # k(-t) = conj(k(t))
    window = np.cos(0.5 * np.pi * t_full / t_max)**2
    integrand = window * lineshape(t_full) * k_full
    rate = 2.0 * t_max * np.fft.ifft(np.fft.ifftshift(integrand))
    rate = rate * (1.0e15 / units.autime_to_fs)                                   # atomic units of time^-1 -> s^-1
    energies = np.fft.fftfreq(m) * m / (2.0 * t_max) * 2.0 * np.pi
    return np.fft.fftshift(energies), np.fft.fftshift(rate)


def calculate_rates(task):
    """the `"task": "rates"` entry of a `semi` JSON file (cli.py:519-570): reads `correlations`, adds `broadening`, `hwhmG`,
    `hwhmL`, `energies`, `ic_rate` (non-negative energies, real part, times the reference's factor 2 pi) and writes `rates`"""
    hwhmG = task.get('hwhmG_ev', 0.01)
    hwhmL = task.get('hwhmL_ev', 1.0e-6)
    sigma = hwhmG / np.sqrt(2.0 * np.log(2.0)) / units.hartree_to_ev
    gamma = hwhmL / units.hartree_to_ev
    broad = task.get('broadening', 'gaussian')
    if broad == "gaussian":
        lineshape = gaussian(sigma)
    elif broad == "lorentzian":
        lineshape = lorentzian(gamma)
    elif broad == "voigtian":
        lineshape = voigtian(sigma, gamma)
    else:
        raise ValueError("'broadening' should be one of 'gaussian', 'lorentzian' or 'voigtian'")
    corr_file = task.get('correlations', 'correlations.npz')
    rate_file = task.get('rates', 'correlations.npz')
    logger.info(f"compute rates from correlation functions in '{corr_file}'")
    data = dict(np.load(corr_file))
    energies, ic_rate = rate_from_correlation(data['times'], data['ic_correlation'], lineshape)
    ic_rate = ic_rate * 2.0 * np.pi
    keep = energies >= 0.0
    data.update(broadening=broad, hwhmG=hwhmG, hwhmL=hwhmL, energies=energies[keep], ic_rate=ic_rate[keep].real)
    logger.info(f"rates are saved to '{rate_file}'")
    np.savez(rate_file, **data)
    return data
