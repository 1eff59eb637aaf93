"""
Data-parallel propagation over the GPUs of one node.

Trajectories never interact on the hot path: rank r owns the contiguous slice [r n/R, (r+1) n/R) of the ensemble
and every rank normalises by the GLOBAL ensemble size (the N of 1/(N probi (2 pi hbar)^d), propagators.py:837,
909), so the combine is ONE all-reduce(SUM) of the (nsteps, 5) correlation buffer per launch interval -- the same
arithmetic as the reference's running average over repetitions, (n_new C_new + n_old C_old)/n_tot
(cli.py:453-458).  One process per GPU, torch.distributed (NCCL on GPUs; gloo in the CPU tests of this logic).
"""
import numpy as np
import torch


def shard_bounds(ntraj, rank, world):
    """[lo, hi) of rank's contiguous slice; sizes differ by at most one and add up to ntraj"""
    assert 0 <= rank < world
    return rank * ntraj // world, (rank + 1) * ntraj // world


def shard_ensemble(zi, probi, rank, world):
    """slice of an ensemble zi (2 dim, n), probi (n,) owned by `rank`"""
    lo, hi = shard_bounds(int(probi.shape[0]), rank, world)
    return zi[:, lo:hi], probi[lo:hi]


def allreduce_rows(rows, n_local, n_total, group=None):
    """
    rows: (nsteps, 5) tensor [Re C, Im C, Re k, Im k, <T+V>_local]; correlation columns are already divided by
    the global N, the energy column is the mean over the LOCAL shard.  In place: sums columns 0..3 over the
    ranks and turns column 4 into the mean over the global ensemble.  One collective.
    """
    import torch.distributed as dist
    rows[:, 4] *= float(n_local) / float(n_total)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=group)
    return rows


def rows_to_correlations(rows, times, energy0_es, hbar=1.0):
    """(autocorrelation, ic_correlation) with the dynamical phase e^{i t E0 / hbar} (propagators.py:841, 906)"""
    rows = rows.detach().cpu().numpy() if isinstance(rows, torch.Tensor) else np.asarray(rows)
    phase = np.exp(1j / hbar * np.asarray(times) * energy0_es)
    return (rows[:, 0] + 1j * rows[:, 1]) * phase, (rows[:, 2] + 1j * rows[:, 3]) * phase


def agree_on_error(err, device='cpu', group=None):
    """
    collective error handling of the task driver: a failure on ONE rank (file validation, empty shard, energy guard, NaN
    guard) raises on EVERY rank before the next collective instead of leaving the others blocked in it.  `err` is the
    exception this rank caught (or None); the failing rank re-raises its own exception, the others a RuntimeError.
    """
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if err is not None:
            raise err
        return
    flag = torch.tensor([0.0 if err is None else 1.0], dtype=torch.float64, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if float(flag.item()) > 0.0:
        if err is not None:
            raise err
        raise RuntimeError("semi dynamics: another rank failed (see its log); stopping all ranks")
