"""C2 throughput: Walton-Manolopoulos on the 5-mode AS model, K fused steps per launch.  usage: wm_probe.py [ntraj] [nsteps]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import potentials, propagators, workloads
T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
m = workloads.as_5modes(0.02)
G = np.diag(m.omega)
pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
pr = propagators.WaltonManolopoulosPropagator(T(G), T(G), 500, 500, device="cuda:0")
torch.manual_seed(0)
pr.initial_conditions(T(m.q0), T(m.p0), T(G), ntraj=n)
dt = workloads.test_time_grid()[0]
pr.propagate(pot, dt, K, m.en_zpt)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pr.propagate(pot, dt, K, m.en_zpt); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"workload": "C2 AS 5 modes WM alpha=beta=500", "ntraj": n, "steps": K, "ms": ms, "traj_steps_per_s": n * K / ms * 1e3,
                  "kernel": pr.kernel_name()}))
