"""C5 throughput: HK on the synthetic sGDML model of the coumarin fixture's shapes (N = 17, d = 51, 200 training points)
through the dense column pipeline; prints per-kernel times.  usage: c5_probe.py [ntraj] [nsteps]"""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import _native, potentials, propagators, workloads
T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
model, pos = workloads.gdml_synthetic()
d = len(pos)
masses = np.full(d, 12.0 * 1822.888486192)
pot = potentials.MolecularGDMLPotential.from_arrays(model, masses, 1.0e-3 * np.ones(d))
G = np.diag(np.full(d, 20.0))
pr = propagators.HermanKlukPropagator(T(G), T(G), device="cuda:0")
torch.manual_seed(0)
pr.initial_conditions(T(pos), T(np.zeros(d)), T(G), ntraj=n)
pr.propagate(pot, 0.5, K, 0.0)
L = _native.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
L.sc_engine_set_timing(pr._engine, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pr.propagate(pot, 0.5, K, 0.0); e1.record(); torch.cuda.synchronize()
kt = np.zeros(8)
L.sc_engine_get_timing_slots(pr._engine, kt.ctypes.data, 8)
ms = e0.elapsed_time(e1)
print(json.dumps({"workload": "C5 synthetic sGDML N=17 d=51 M=200, HK, diagonal widths", "ntraj": n, "steps": K, "ms": ms,
                  "traj_steps_per_s": n * K / ms * 1e3, "gdml_evals_per_s": 4 * n * K / (kt[0] * 1e-3) if kt[0] > 0 else None,
                  "kernel": pr.kernel_name(),
                  "kernel_ms": {"path(sGDML evals)": kt[0], "rk4": kt[1], "rmult": kt[4], "lu": kt[2], "finish": kt[3]}}))
