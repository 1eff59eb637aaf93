"""Walton-Manolopoulos at larger d (AS model): parity against the C oracle on a small ensemble + throughput.
usage: wm_big_probe.py [d] [ntraj_timing] [steps]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import potentials, propagators, workloads
from oracle import oracle
T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
d = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
m = workloads.as_synthetic(d, 0.02)
G = np.diag(m.omega)
dt, _ = workloads.test_time_grid()
zi, probi = oracle.sample_ensemble(G, G, m.q0, m.p0, 48, np.random.default_rng(5))
nt = 6
ref = oracle.run(oracle.Potential.morse(m.omega, m.chi, m.nac), oracle.Consts(G, G, G, m.q0, m.p0, alpha=500.0, beta=500.0), zi, probi,
                 dt, nt, m.en_zpt, wm=True)
pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
pr = propagators.WaltonManolopoulosPropagator(T(G), T(G), 500, 500, device="cuda:0")
pr.set_ensemble(T(m.q0), T(m.p0), T(G), T(zi), T(probi))
auto = [pr.autocorrelation(m.en_zpt)]
a, i = pr.propagate(pot, dt, nt - 1, m.en_zpt)
auto.extend(a)
auto = np.array(auto); r = np.asarray(ref['autocorrelation'])
err = np.max(np.abs(auto - r)) / np.max(np.abs(r))
print("parity d=%d: rel err %.2e  kernel %s" % (d, err, pr.kernel_name()))
torch.manual_seed(0)
pr.initial_conditions(T(m.q0), T(m.p0), T(G), ntraj=n)
pr.propagate(pot, dt, K, m.en_zpt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pr.propagate(pot, dt, K, m.en_zpt); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"workload": "AS %d modes WM alpha=beta=500" % d, "ntraj": n, "steps": K, "ms": ms, "traj_steps_per_s": n * K / ms * 1e3,
                  "kernel": pr.kernel_name()}))
