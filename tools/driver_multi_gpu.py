"""`semi dynamics` driver sharded over the ranks of a torchrun launch (one process per GPU, NCCL): the injected
reference ensemble of the 5-mode AS fixture must reproduce the reference's correlation functions (golden) at 1e-9.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/driver_multi_gpu.py"""
import os, sys, tempfile
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
torch.set_default_dtype(torch.float64)
from semiclassical_b200 import dynamics, units, workloads
import helpers
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
g = helpers.load_golden("hk_as5_chi002")
tmp = tempfile.mkdtemp() if rank == 0 else None
box = [tmp]
dist.broadcast_object_list(box, src=0)
tmp = box[0]
model_file = os.path.join(tmp, "AS_model.dat")
if rank == 0:
    rows = workloads._AS5_ROWS
    np.savetxt(model_file, np.column_stack((rows, np.full(len(rows), 0.02))), fmt="%.10f")
dist.barrier()
nt, n = int(g['nt']), len(g['probi'])
out = os.path.join(tmp, "correlations.npz")
task = {"task": "dynamics", "potential": {"type": "anharmonic AS", "model_file": model_file}, "propagator": "HK",
        "batch_size": n, "num_trajectories": n, "num_steps": nt, "time_step_fs": float(g['dt']) * units.autime_to_fs,
        "results": {"correlations": out}, "calc_norm_every": 40}
import logging
norms = []
class _Grab(logging.Handler):
    def emit(self, record):
        m = record.getMessage()
        if "norm=" in m:
            norms.append(float(m.split("norm=")[1]))
logging.getLogger("semiclassical_b200.dynamics").addHandler(_Grab())
logging.getLogger("semiclassical_b200.dynamics").setLevel(logging.INFO)
dynamics.run_semiclassical_dynamics(task, device=f"cuda:{torch.cuda.current_device()}", ensembles=[(g['zi'], g['probi'])],
                                    steps_per_launch=23)
dist.barrier()
if rank == 0:
    data = dict(np.load(out))
    ea, ei = helpers.relerr(data['autocorrelation'], g['autocorrelation']), helpers.relerr(data['ic_correlation'], g['ic_correlation'])
    print(f"driver on {world} ranks: max rel err autocorrelation {ea:.2e}, ic_correlation {ei:.2e}, trajectories {int(data['trajectories'])}")
    assert ea < 1e-9 and ei < 1e-9
# the sharded norm (all-gather of the ket vectors + blocks + all-reduce) is the norm of the GLOBAL wavefunction on every rank:
# compare with the single-device norm of the whole ensemble
from semiclassical_b200 import potentials, propagators
pot = helpers.potential_from_golden(g)
pr = helpers.propagator_from_golden(g, f"cuda:{torch.cuda.current_device()}")
ref_norms = [pr.norm()]
for t0 in (40, 80):
    pr.propagate(pot, float(g['dt']), 40, float(g['energy0_es']))
    ref_norms.append(pr.norm())
err = max(abs(a / b - 1.0) for a, b in zip(norms, ref_norms))
print(f"rank {rank}: driver-logged sharded norms {norms} vs single-device {ref_norms}: max rel diff {err:.2e} (6 logged decimals)")
assert len(norms) == 3 and err < 2e-6
# full precision: a sharded propagator next to the whole-ensemble one
from semiclassical_b200 import distributed
lo, hi = distributed.shard_bounds(n, rank, world)
prs = helpers.propagator_from_golden(g, f"cuda:{torch.cuda.current_device()}", nslice=slice(lo, hi))
prs.propagate(pot, float(g['dt']), 80, float(g['energy0_es']), group=True)
ns = prs.norm(group=True)
print(f"rank {rank}: sharded norm {ns!r} vs single-device {ref_norms[-1]!r}: rel diff {abs(ns / ref_norms[-1] - 1.0):.2e}")
assert abs(ns / ref_norms[-1] - 1.0) < 1e-12
dist.barrier()
dist.destroy_process_group()
