"""shared helpers of the parity tests: build engine objects from golden fixtures / workloads"""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def T(x):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))


def potential_from_golden(g):
    from semiclassical_b200 import potentials
    kind = str(g['potential'])
    if kind == "morse":
        return potentials.MorsePotential(T(g['omega']), T(g['chi']).clone(), T(g['nac']))
    if kind == "rotated_morse":
        return potentials.RotatedMorsePotential(T(g['omega']), T(g['chi']).clone(), T(g['nac']), T(g['Q']))
    if kind == "nonharmonic":
        return potentials.NonHarmonicPotential(T(g['eps']), T(g['b']))
    if kind == "harmonic":
        return potentials.MolecularHarmonicPotential.from_arrays(g['pos0'], g['energy0'], g['grad0'], g['hess0'],
                                                                 g['masses'], g['nac'], float(g['origin']))
    if kind == "gdml":
        if 'gdml_model_fixture' in g.files:              # the model arrays live in another fixture
            gm = load_golden(str(g['gdml_model_fixture']))
            g = dict(g.items())
            g.update({k: gm[k] for k in gm.files if k.startswith('gdml_')})
        D = g['gdml_R_desc'].shape[0]
        n_atoms = len(g['masses']) // 3
        model = dict(sig=int(g['gdml_sig']), c=float(g['gdml_c']), std=float(g['gdml_std']), R_desc=g['gdml_R_desc'],
                     R_d_desc_alpha=g['gdml_R_d_desc_alpha'], perms=np.arange(n_atoms)[None, :],
                     tril_perms_lin=np.arange(D))
        return potentials.MolecularGDMLPotential.from_arrays(model, g['masses'], g['nac'], float(g['origin']))
    raise ValueError(kind)


def propagator_from_golden(g, device, nslice=None):
    from semiclassical_b200 import propagators
    if str(g['kind']) == "WM":
        pr = propagators.WaltonManolopoulosPropagator(T(g['Gamma_i']), T(g['Gamma_t']), float(g['alpha']), float(g['beta']),
                                                      device=device)
    else:
        pr = propagators.HermanKlukPropagator(T(g['Gamma_i']), T(g['Gamma_t']), device=device)
    zi, probi = g['zi'], g['probi']
    ntot = len(probi)
    if nslice is not None:
        zi, probi = zi[:, nslice], probi[nslice]
    pr.set_ensemble(T(g['q0']), T(g['p0']), T(g['Gamma_0']), T(zi), T(probi), ntraj_total=ntot)
    return pr


def relerr(a, b):
    """max_t |a - b| / max_t |b|  (the parity metric of SURVEY.md section 7.2-4)"""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())


def regenerate_ensemble(g):
    """ensemble of a fixture that stores only its numpy seed (oracle.sample_ensemble stream) and a checksum"""
    from oracle import oracle
    zi, probi = oracle.sample_ensemble(g['Gamma_i'], g['Gamma_0'], g['q0'], g['p0'], int(g['ntraj']),
                                       np.random.default_rng(int(g['ensemble_seed'])))
    chk = np.array([zi.sum(), np.abs(zi).sum(), zi[0, 0], zi[-1, -1]])
    assert np.allclose(chk, g['zi_checksum'], rtol=1e-13, atol=0.0), "numpy's PCG64 normal stream changed: regenerate the fixture"
    return zi, probi
