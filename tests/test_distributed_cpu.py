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


def _agree_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seen = []
        # round 1: nobody failed -> nobody raises; round 2: rank 1 failed -> BOTH ranks raise before the next collective
        distributed.agree_on_error(None)
        seen.append("ok")
        try:
            distributed.agree_on_error(ValueError("energy guard on rank 1") if rank == 1 else None)
            seen.append("no exception")
        except ValueError as err:
            seen.append("own:" + str(err))
        except RuntimeError as err:
            seen.append("peer:" + str(err)[:40])
        # the ranks are still in lock step: one more collective completes
        t = torch.ones(1)
        dist.all_reduce(t)
        seen.append(int(t.item()))
        with open(out + str(rank), "w") as f:
            f.write(repr(seen))
    finally:
        dist.destroy_process_group()


def test_error_on_one_rank_raises_on_all(tmp_path):
    """a rank that raises (validation, energy guard) must not leave the others blocked in the next all-reduce"""
    out = str(tmp_path / "seen")
    mp.spawn(_agree_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    s0, s1 = (eval(open(out + str(r)).read()) for r in (0, 1))
    assert s0[0] == "ok" and s1[0] == "ok"
    assert s0[1].startswith("peer:") and s1[1] == "own:energy guard on rank 1"
    assert s0[2] == 2 and s1[2] == 2
