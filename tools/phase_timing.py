"""builds an instrumented copy of the library (-DSC_PHASE_TIMING) and prints the cycle split of k_hk_mma"""
import ctypes, os, subprocess, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from semiclassical_b200 import _native
lib_dbg = os.path.join(_native.HERE, "lib", "libsemiclassical_b200_timing.so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + _native.NVCC_FLAGS + ["-DSC_PHASE_TIMING", "-o", lib_dbg, os.path.join(_native.CSRC, "sc_engine.cu")])
_native.LIB_PATH = lib_dbg
_native.needs_build = lambda: False
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import workloads, potentials, propagators
T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
n, K = int(sys.argv[1]) if len(sys.argv) > 1 else 1480, int(sys.argv[2]) if len(sys.argv) > 2 else 10
dense = len(sys.argv) > 3 and sys.argv[3] == "dense"
d = 60
m = workloads.as_synthetic(d)
G = np.diag(m.omega); q0, p0 = m.q0, m.p0
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
pr.propagate(pot, dt, 2, m.en_zpt)
L = _native.lib()
buf = (ctypes.c_ulonglong * 16)()
L.sc_debug_phase_cycles(buf, 1)
pr.propagate(pot, dt, K, m.en_zpt)
L.sc_debug_phase_cycles(buf, 1)
names = ["load", "potential", "gemm(4 stages)", "elementwise(4)", "prefactor assembly", "LU", "corr+reduce", "restore Us", "store"]
ntr = (n + 147) // 148   # trajectories of CTA 0
if pr.kernel_name().startswith("k_rk4_chunk"):
    ntr = n * 3 / 444.0   # (trajectory, chunk) items of CTA 0
    print("chunked path: cycles per CHUNK-step of CTA 0 (3 CTAs per SM run concurrently)")
tot = sum(buf[:9])
print(f"CTA 0: {ntr} trajectories x {K} steps; cycles per trajectory-step = {tot/(ntr*K):.0f}")
for i, nm in enumerate(names):
    print(f"  {nm:20s} {buf[i]/(ntr*K):10.0f} cycles/traj-step  {100*buf[i]/tot:5.1f}%")
