"""throughput of the all-pairs norm() kernel (k_gauss_sum, sc_gauss.cuh): pairs/s and FP64 TFLOP/s of the two bilinear
contractions (algorithmic flops per pair: 2 dot products of length 2d = 8 d), synthetic 60-mode AS ensemble"""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import workloads, potentials, propagators
T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
d = int(sys.argv[1]) if len(sys.argv) > 1 else 60
m = workloads.as_synthetic(d)
G = np.diag(m.omega)
pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
dt = workloads.test_time_grid()[0]
out = []
for n in (9472, 37888, 151552):
    pr = propagators.HermanKlukPropagator(T(G), T(G), device='cuda:0')
    torch.manual_seed(0)
    pr.initial_conditions(T(m.q0), T(m.p0), T(G), ntraj=n)
    pr.propagate(pot, dt, 5, m.en_zpt)
    pr.norm()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); nrm = pr.norm(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out.append({"ntraj": n, "dim": d, "norm": nrm, "ms": ms, "pairs_per_s": n * n / ms * 1e3,
                "tflops_contraction": 8.0 * d * n * n / ms * 1e3 / 1e12})
    print(json.dumps(out[-1]), flush=True)
    del pr
