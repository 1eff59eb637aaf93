"""the reference's driver loop {autocorrelation; ic_correlation; step} (cli.py:401-436) against one fused propagate() call:
AS model, d modes, n trajectories.  usage: steploop_probe.py [d] [ntraj] [nsteps]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import potentials, propagators, workloads
T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
d = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n = int(sys.argv[2]) if len(sys.argv) > 2 else 148000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
m = workloads.as_synthetic(d, 0.02) if d != 5 else workloads.as_5modes(0.02)
G = np.diag(m.omega)
pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
pr = propagators.HermanKlukPropagator(T(G), T(G), device="cuda:0")
torch.manual_seed(0)
pr.initial_conditions(T(m.q0), T(m.p0), T(G), ntraj=n)
dt = workloads.test_time_grid()[0]
pr.propagate(pot, dt, K, m.en_zpt)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr.propagate(pot, dt, K, m.en_zpt)
torch.cuda.synchronize()
t_fused = time.perf_counter() - t0
for _ in range(2):
    pr.autocorrelation(m.en_zpt); pr.ic_correlation(pot, m.en_zpt); pr.step(pot, dt)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(K):
    pr.autocorrelation(m.en_zpt); pr.ic_correlation(pot, m.en_zpt); pr.step(pot, dt)
torch.cuda.synchronize()
t_loop = time.perf_counter() - t0
print(json.dumps({"d": d, "ntraj": n, "steps": K, "fused_traj_steps_per_s": n * K / t_fused, "step_loop_traj_steps_per_s": n * K / t_loop,
                  "loop_over_fused": t_loop / t_fused, "ms_per_loop_iteration": 1e3 * t_loop / K, "kernel": pr.kernel_name()}))
