"""small driver for ncu: one fused launch of the HK step kernel on the 60-mode AS model"""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import workloads, potentials, propagators
T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dense = len(sys.argv) > 3 and sys.argv[3] == "dense"
d = int(sys.argv[4]) if len(sys.argv) > 4 else 60
m = workloads.as_synthetic(d)
G = np.diag(m.omega)
q0, p0 = m.q0, m.p0
if dense:
    Q = workloads.random_orthogonal(d, 11)
    G = Q @ G @ Q.T; G = 0.5 * (G + G.T); q0 = Q @ q0; p0 = Q @ p0
    pot = potentials.RotatedMorsePotential(T(m.omega), T(m.chi), T(m.nac), T(Q))
else:
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
pr = propagators.HermanKlukPropagator(T(G), T(G), device='cuda:0')
torch.manual_seed(0)
pr.initial_conditions(T(q0), T(p0), T(G), ntraj=n)
dt = workloads.test_time_grid()[0]
pr.propagate(pot, dt, 1, m.en_zpt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
a, i = pr.propagate(pot, dt, K, m.en_zpt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"kernel={pr.kernel_name()} n={n} K={K} {ms:.3f} ms  {n*K/ms*1e3:.4g} traj-steps/s  C={a[-1]:.6f}")
