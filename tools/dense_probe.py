"""throughput of the dense column pipeline (sc_stream.cuh) on a synthetic harmonic 'molecule': dense Hessian,
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
m = workloads.harmonic_molecule_synthetic(d)
G0, q0 = m['Gamma_0'], m['q0']
pot = potentials.MolecularHarmonicPotential.from_arrays(m['pos0'], m['energy0'], m['grad0'], m['hess0'], m['masses'], m['nac'])
pr = propagators.HermanKlukPropagator(T(G0), T(G0), device="cuda:0")
torch.manual_seed(0)
pr.initial_conditions(T(q0), T(np.zeros(d)), T(G0), ntraj=n)
dt = workloads.test_time_grid()[0]
a, i = pr.propagate(pot, dt, K, m['en_zpt'])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); a, i = pr.propagate(pot, dt, K, m['en_zpt']); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
if os.environ.get("SC_PROBE_SLOTS"):
    from semiclassical_b200 import _native
    kt = np.zeros(8)
    _native.lib().sc_engine_set_timing(pr._engine, 1)
    pr.propagate(pot, dt, K, m['en_zpt'])
    torch.cuda.synchronize()
    _native.lib().sc_engine_get_timing_slots(pr._engine, kt.ctypes.data, 8)
    print("kernel_ms of the last launch: path %.2f rk4 %.2f rmult %.2f lu %.2f finish %.2f hess %.2f" % (kt[0], kt[1], kt[4], kt[2], kt[3], kt[5]))
print(json.dumps({"workload": f"harmonic molecule-like, d={d}, d'={d-6}, dense Gamma", "ntraj": n, "steps": K, "ms": ms,
                  "traj_steps_per_s": n * K / ms * 1e3, "kernel": pr.kernel_name(), "C_last": [a[-1].real, a[-1].imag]}))
