"""
CPU tests (gloo, world_size 2) of the host logic of the N > 1 path: contiguous sharding of the ensemble and the
single all-reduce of the per-step correlation rows (semiclassical_b200/distributed.py).  Each rank computes its
shard's rows with the CPU oracle (the checker -- the GPU ranks get the same rows from sc_engine_step_dev), the
combination must reproduce the unsharded run and the reference golden.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from semiclassical_b200 import distributed


def test_shard_bounds_partition():
    for n in (1, 7, 1000, 10**6):
        for world in (1, 2, 3, 8):
            b = [distributed.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, nt, out):
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = helpers.load_golden(name)
        pot, consts, wm = oracle.from_golden(g)
        n = len(g['probi'])
        zi, probi = distributed.shard_ensemble(g['zi'], g['probi'], rank, world)
        dt, e0 = float(g['dt']), float(g['energy0_es'])
        r = oracle.run(pot, consts, np.ascontiguousarray(zi), np.ascontiguousarray(probi), dt, nt, e0, wm=wm, ntraj_norm=n,
                       nthreads=2)
        times = np.cumsum(np.concatenate(([0.0], np.full(nt - 1, dt))))
        phase = np.exp(1j * times * e0)
        a, k = r['autocorrelation'] / phase, r['ic_correlation'] / phase
        rows = torch.from_numpy(np.stack((a.real, a.imag, k.real, k.imag, r['energy']), axis=1).copy())
        distributed.allreduce_rows(rows, len(probi), n)
        auto, ic = distributed.rows_to_correlations(rows, times, e0)
        if rank == 0:
            np.savez(out, auto=auto, ic=ic, energy=rows[:, 4].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["hk_as5_rot", "wm_as5_chi002"])
def test_two_ranks_reproduce_the_unsharded_run(name, tmp_path):
    from oracle import oracle
    nt, world = 12, 2
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(world, _free_port(), name, nt, out), nprocs=world, join=True)
    res = np.load(out)
    g = helpers.load_golden(name)
    assert helpers.relerr(res['auto'], g['autocorrelation'][:nt]) < 1e-12
    assert helpers.relerr(res['ic'], g['ic_correlation'][:nt]) < 1e-12
    pot, consts, wm = oracle.from_golden(g)
    full = oracle.run(pot, consts, g['zi'], g['probi'], float(g['dt']), nt, float(g['energy0_es']), wm=wm)
    assert np.abs(res['energy'] - full['energy']).max() < 1e-12 * max(1.0, np.abs(full['energy']).max())
