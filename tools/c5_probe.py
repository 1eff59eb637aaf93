import os, sys
import numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import workloads, potentials, propagators
import helpers
T = helpers.T
model, pos = workloads.gdml_synthetic()
d = len(pos)
masses = np.full(d, 12.0 * 1822.888486192)
potg = potentials.MolecularGDMLPotential.from_arrays(model, masses, 1.0e-3 * np.ones(d))
Gg = np.diag(np.full(d, 20.0))
pr = propagators.HermanKlukPropagator(T(Gg), T(Gg), device="cuda:0")
torch.manual_seed(0)
pr.initial_conditions(T(pos), T(np.zeros(d)), T(Gg), ntraj=20000)
pr.propagate(potg, 0.5, 3, 0.0)
torch.cuda.synchronize()
