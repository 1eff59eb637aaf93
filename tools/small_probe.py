"""k_hk_small on the two small BASELINE configs: C1 (AS, 5 modes, Morse, diagonal widths) at 10^6 trajectories and C3
(methylium, 12 modes, harmonic, dense rank-6 widths) at 10^5; usage: small_probe.py [c1|c3] [ntraj] [steps]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import workloads, potentials, propagators
import helpers
T = helpers.T
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (1000000 if which == "c1" else 100000)
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
if which == "c1":
    dt, _ = workloads.test_time_grid()
    m = workloads.as_5modes(0.02)
    G = np.diag(m.omega)
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    pr = propagators.HermanKlukPropagator(T(G), T(G), device="cuda:0")
    q0, p0, G0, e0 = m.q0, m.p0, G, m.en_zpt
else:
    g = helpers.load_golden("hk_methylium")
    pot = helpers.potential_from_golden(g)
    pr = propagators.HermanKlukPropagator(T(g['Gamma_i']), T(g['Gamma_t']), device="cuda:0")
    q0, p0, G0, e0, dt = g['q0'], g['p0'], g['Gamma_0'], float(g['energy0_es']), float(g['dt'])
torch.manual_seed(0)
pr.initial_conditions(T(q0), T(p0), T(G0), ntraj=n)
pr.propagate(pot, dt, K, e0)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    auto, _ic = pr.propagate(pot, dt, K, e0)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
d = pr.dim
flops = n * K * (4 * 2 * d * d * 2 * d if which == "c3" else 4 * 2 * d * 2 * d)     # H U products only
print(json.dumps({"workload": which, "ntraj": n, "dim": d, "steps": K, "ms": ms, "traj_steps_per_s": n * K / ms * 1e3,
                  "kernel": pr.kernel_name(), "hu_product_tflops": flops / ms * 1e-9, "C_last": [auto[-1].real, auto[-1].imag]}))
