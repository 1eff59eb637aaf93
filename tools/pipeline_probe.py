"""per-kernel device times of the column pipeline (k_qp_path, k_rk4_wcols, k_lu_mma) from the engine's CUDA-event timing
usage: pipeline_probe.py NTRAJ NSTEPS   (env: SC_CHUNK_K, SC_CHUNK_CTAS, SC_LU_CTAS, SC_WCOLS_TILES, SC_LU_DFMA)"""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, '/root/repo')
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import workloads, potentials, propagators, _native
T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
n, K = int(sys.argv[1]), int(sys.argv[2])
m = workloads.as_synthetic(60)
G = np.diag(m.omega)
pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
pr = propagators.HermanKlukPropagator(T(G), T(G), device='cuda:0')
torch.manual_seed(0)
pr.initial_conditions(T(m.q0), T(m.p0), T(G), ntraj=n)
dt = workloads.test_time_grid()[0]
pr.propagate(pot, dt, K, m.en_zpt)
L = _native.lib()
L.sc_engine_set_timing(pr._engine, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pr.propagate(pot, dt, K, m.en_zpt)
e1.record(); torch.cuda.synchronize()
kt = np.zeros(4)
L.sc_engine_get_timing(pr._engine, kt.ctypes.data)
wall = e0.elapsed_time(e1)
print(f"CTAS={os.environ.get('SC_CHUNK_CTAS')} LU={os.environ.get('SC_LU_CTAS')} K={os.environ.get('SC_CHUNK_K')}: wall {wall:.1f} ms, rk4 {kt[1]:.1f} ms, lu {kt[2]:.1f} ms, qp {kt[0]:.1f}  -> {n*K/wall*1e3:.4g} traj-steps/s")
