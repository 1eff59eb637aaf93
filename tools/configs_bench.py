"""throughput of the engine on the other BASELINE.json configs (C1, C2, C3, C5: parity-test cases, not the headline):
trajectory-steps/s through the public API with the state resident on the GPU, ensembles sampled on the device.
C3 uses the arrays of the methylium golden fixture (fchk files are not in this repo), C5 the synthetic sGDML model
of workloads.gdml_synthetic (N = 17 atoms, 200 training points) at the 4-atom fixture's width matrices scaled up."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import workloads, potentials, propagators
import helpers
T = helpers.T
dev = "cuda:0"


def run(tag, pr, pot, q0, p0, G0, n, dt, K, e0, reps=3):
    torch.manual_seed(0)
    pr.initial_conditions(T(q0), T(p0), T(G0), ntraj=n)
    pr.propagate(pot, dt, K, e0)
    torch.cuda.synchronize()
    e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0_.record()
    for _ in range(reps):
        pr.propagate(pot, dt, K, e0)
    e1_.record(); torch.cuda.synchronize()
    ms = e0_.elapsed_time(e1_) / reps
    line = {"config": tag, "ntraj": n, "dim": pr.dim, "steps_per_launch": K, "ms": ms, "traj_steps_per_s": n * K / ms * 1e3,
            "kernel": pr.kernel_name()}
    print(json.dumps(line), flush=True)


dt5, _ = workloads.test_time_grid()
m = workloads.as_5modes(0.02)
G = np.diag(m.omega)
pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
run("C1 AS 5 modes HK", propagators.HermanKlukPropagator(T(G), T(G), device=dev), pot, m.q0, m.p0, G, 1000, dt5, 50, m.en_zpt)
run("C1 AS 5 modes HK (10^6 trajectories)", propagators.HermanKlukPropagator(T(G), T(G), device=dev), pot, m.q0, m.p0, G, 1000000, dt5, 50, m.en_zpt)
run("C2 AS 5 modes WM", propagators.WaltonManolopoulosPropagator(T(G), T(G), 500, 500, device=dev), pot, m.q0, m.p0, G, 10000, dt5, 50, m.en_zpt)
run("C2 AS 5 modes WM (10^6 trajectories)", propagators.WaltonManolopoulosPropagator(T(G), T(G), 500, 500, device=dev), pot, m.q0, m.p0, G, 1000000, dt5, 20, m.en_zpt)
g = helpers.load_golden("hk_methylium")
potm = helpers.potential_from_golden(g)
run("C3 methylium harmonic HK", propagators.HermanKlukPropagator(T(g['Gamma_i']), T(g['Gamma_t']), device=dev), potm, g['q0'], g['p0'],
    g['Gamma_0'], 100000, float(g['dt']), 100, float(g['energy0_es']))
model, pos = workloads.gdml_synthetic()
d = len(pos)
masses = np.full(d, 12.0 * 1822.888486192)
potg = potentials.MolecularGDMLPotential.from_arrays(model, masses, 1.0e-3 * np.ones(d))
Gg = np.diag(np.full(d, 20.0))
run("C5 sGDML N=17 HK", propagators.HermanKlukPropagator(T(Gg), T(Gg), device=dev), potg, pos, np.zeros(d), Gg, 20000, 0.5, 5,
    0.0, reps=2)
