"""throughput of the general path (k_hk_mma split + batched LU) on a synthetic harmonic 'molecule': dense Hessian,
dense width matrices with 6 zero modes (d' = d - 6), the shape of the reference's molecular use case (C3 at size d)
usage: dense_probe.py [d] [ntraj] [nsteps]"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import potentials, propagators, workloads
T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
d = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n = int(sys.argv[2]) if len(sys.argv) > 2 else 29600
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
rng = np.random.default_rng(3)
V, _ = np.linalg.qr(rng.standard_normal((d, d)))
w = np.concatenate((np.zeros(6), np.linspace(200.0, 3400.0, d - 6) / 219474.63))     # 6 zero modes, masses = 1
G0 = (V * w[None, :]) @ V.T
G0 = 0.5 * (G0 + G0.T)
hess = (V * (w * w)[None, :]) @ V.T
hess = 0.5 * (hess + hess.T)
pos0 = rng.standard_normal(d)
q0 = pos0 + V[:, 6:] @ (0.3 * rng.standard_normal(d - 6) / np.sqrt(w[6:]))      # displaced along the vibrations only
pot = potentials.MolecularHarmonicPotential.from_arrays(pos0, 0.0, np.zeros(d), hess, np.ones(d), 1.0e-3 * rng.standard_normal(d))
pr = propagators.HermanKlukPropagator(T(G0), T(G0), device="cuda:0")
torch.manual_seed(0)
pr.initial_conditions(T(q0), T(np.zeros(d)), T(G0), ntraj=n)
dt = workloads.test_time_grid()[0]
a, i = pr.propagate(pot, dt, K, 0.5 * w.sum())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); a, i = pr.propagate(pot, dt, K, 0.5 * w.sum()); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"workload": f"harmonic molecule-like, d={d}, d'={d-6}, dense Gamma", "ntraj": n, "steps": K, "ms": ms,
                  "traj_steps_per_s": n * K / ms * 1e3, "kernel": pr.kernel_name(), "C_last": [a[-1].real, a[-1].imag]}))
