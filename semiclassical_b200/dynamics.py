"""
Driver of the `dynamics` task of `semi` (cli.run_semiclassical_dynamics, cli.py:171-476) on top of the B200 engine.

Same JSON task, same `correlations.npz` (keys, running average over repetitions, `overwrite=false` accumulation, the
`times = linspace(0, nt dt, nt)` grid of cli.py:310-311 whose spacing is not dt, NaN guard, C(0) = 1 check), but the
time loop {autocorrelation; ic_correlation; step} x nt of one repetition (cli.py:401-436, two host syncs per step)
becomes a few fused launches of `steps_per_launch` time steps, and a repetition may be as large as the GPU memory
allows (116 KB per trajectory at 60 modes) instead of the reference's 10^4.  With torch.distributed initialised
(one process per GPU) every repetition is sharded over the ranks and combined by one all-reduce per launch
(semiclassical_b200.distributed); rank 0 writes the file.

Model potentials ("anharmonic AS") need nothing else.  Molecular potentials ("harmonic", "gdml") read Gaussian
formatted checkpoint files: pass the reference's `semiclassical.readers` module (or anything with the same
`FormattedCheckpointFile` interface, readers.py) as `readers=`; parsing fchk files is setup-time host work and out
of the scope of this package.
"""
import logging
import os

import numpy as np
import torch

from semiclassical_b200 import distributed, potentials, propagators, units
from semiclassical_b200.units import hbar

logger = logging.getLogger(__name__)


class ConfigurationError(Exception):
    pass


def _as_model(model_file):
    """frequencies, displacements, NACs and anharmonicities of an adiabatic-shift model file (cli.py:229-283)"""
    data = torch.from_numpy(np.loadtxt(model_file))
    if len(data.shape) == 1:
        data = torch.reshape(data, (1, -1))
    omega = data[:, 0] / units.hartree_to_wavenumbers
    S = data[:, 1]
    nac = data[:, 2]
    dQ = torch.sqrt(2.0 * abs(S) / omega) * torch.sign(S)
    dQ[omega == 0.0] = 0.0
    chi = data[:, 3]
    potential = potentials.MorsePotential(omega, chi, nac)
    return potential, dQ, 0.0 * dQ, torch.diag(omega), torch.sum(hbar / 2.0 * omega).item()


def _molecular(p, readers):
    if readers is None:
        try:
            from semiclassical import readers  # the reference package, if it is installed next to this one
        except Exception as err:
            raise ConfigurationError(f"potential type '{p['type']}' reads formatted checkpoint files: pass readers= "
                                     f"(semiclassical.readers is not importable: {err})")
    with open(p['coupling']) as f:
        nacs_fchk = readers.FormattedCheckpointFile(f)
    if p['type'] == "harmonic":
        with open(p['ground']) as f:
            freq_fchk = readers.FormattedCheckpointFile(f)
        potential = potentials.MolecularHarmonicPotential(freq_fchk, nacs_fchk)
    else:
        potential = potentials.MolecularGDMLPotential(np.load(p['ground'], allow_pickle=True), nacs_fchk)
    with open(p['excited']) as f:
        excited_fchk = readers.FormattedCheckpointFile(f)
    x0, Gamma_0, en_zpt = excited_fchk.vibrational_groundstate()
    q0 = torch.from_numpy(x0)
    return potential, q0, torch.zeros_like(q0), torch.from_numpy(Gamma_0), en_zpt, excited_fchk


def run_semiclassical_dynamics(task, device='cuda', readers=None, ensembles=None, steps_per_launch=20):
    """
    Parameters
    ----------
    task      : JSON structure of one `"task": "dynamics"` entry (cli.py:171)
    device    : CUDA device (there is no CPU path)
    readers   : module providing FormattedCheckpointFile for molecular potentials
    ensembles : optional list of (zi (2 dim, n), probi (n,)) per repetition injected instead of sampling
                (parity tests: the reference's RNG stream is device dependent)

    Returns the dictionary that was written to the npz file (rank 0; other ranks return the same arrays).
    """
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None and world > 1:
        device = torch.device('cuda', torch.cuda.current_device())

    def agree(err):
        distributed.agree_on_error(err, device)

    def guarded(fn):
        try:
            return fn(), None
        except Exception as err:   # noqa: BLE001 -- re-raised on all ranks by agree()
            return None, err

    p = task['potential']
    excited_fchk = None
    if p['type'] in ("harmonic", "gdml"):
        potential, q0, p0, Gamma_0, en_zpt, excited_fchk = _molecular(p, readers)
    elif p['type'] == "anharmonic AS":
        potential, q0, p0, Gamma_0, en_zpt = _as_model(p['model_file'])
    else:
        raise ConfigurationError(f"Unknown potential type in {task['potential']}")
    if hasattr(potential, "minimize"):
        logger.info("find minimum on final potential energy surface")
        potential.minimize(q0)
    adiabatic_gap = (excited_fchk.total_energy() - potential.total_energy()) if excited_fchk is not None else np.nan
    Gamma_i = Gamma_t = Gamma_0

    dt = task['time_step_fs'] / units.autime_to_fs
    nt = task['num_steps']
    # explicit dtypes: the reference relies on torch.set_default_dtype(float64) in its CLI (cli.py:121); this driver must
    # produce the same float64 grid (and pass the overwrite=False array_equal check) whatever the process-wide default is
    times = torch.linspace(0.0, nt * dt, nt, dtype=torch.float64)
    batch_size = task.get('batch_size', 10000)
    num_trajectories = task.get('num_trajectories', 50000)
    num_repetitions = max(num_trajectories // batch_size, 1)
    num_samples = min(batch_size, num_trajectories)
    propagator_name = task.get('propagator', 'HK')
    if num_samples < world:
        raise ConfigurationError(f"{num_samples} trajectories per repetition cannot be sharded over {world} ranks")
    if world > 1 and not getattr(potential, '_fused_step', True):
        raise ConfigurationError("this potential is evaluated through the stage interface and cannot be sharded over ranks")

    filename = task['results'].get('correlations', 'correlations.npz')

    def prepare_file():
        if task['results'].get('overwrite', True) is True or (not os.path.exists(filename)):
            np.savez(filename, propagator=propagator_name, times=times, autocorrelation=np.zeros((nt,), dtype=complex),
                     ic_correlation=np.zeros((nt,), dtype=complex), adiabatic_gap=adiabatic_gap, zero_point_energy=en_zpt,
                     trajectories=0)
        else:
            assert task.get('manual_seed', None) is None, \
                "Multiple runs with the same sequence of random numbers make no sense! Do not use `manual_seed` and `overwrite=False` at the same time"
            data = np.load(filename)
            assert np.array_equal(data['times'], times.numpy()), \
                f"Time steps in {filename} differ. Delete the old file or change the grid for time propagation."
            assert data['propagator'] == propagator_name, "Data produced with different propagators cannot be added."

    _, err = guarded(prepare_file) if rank == 0 else (None, None)
    agree(err)
    seed = task.get('manual_seed', None)
    if seed is not None:
        logger.warning("The random number generator should not be seeded manually unless for debugging!")
        torch.manual_seed(seed)
    calc_norm_every = task.get('calc_norm_every', 0)
    group = True if world > 1 else None

    data = None
    # one propagator for all repetitions (the reference builds a new one each time, cli.py:376-383): a new ensemble of the
    # same size reuses the engine and its device buffers, and installing it resets time, state and branch trackers
    if propagator_name == "WM":
        alpha = task.get('cell_width', 10000.0)
        propagator = propagators.WaltonManolopoulosPropagator(Gamma_i, Gamma_t, alpha, alpha, device=device)
    else:
        propagator = propagators.HermanKlukPropagator(Gamma_i, Gamma_t, device=device)
    norm_warned = False
    for repetition in range(num_repetitions):
        logger.info(f"*** Repetition {repetition+1} ***")
        lo, hi = distributed.shard_bounds(num_samples, rank, world)

        def install():
            if ensembles is not None:
                zi, probi = ensembles[repetition]
                propagator.set_ensemble(q0, p0, Gamma_0, torch.as_tensor(zi)[:, lo:hi], torch.as_tensor(probi)[lo:hi],
                                        ntraj_total=num_samples)
            else:
                # one 64-bit seed per repetition (rank 0's torch generator, hence `manual_seed`), the same on every rank: the
                # counter-based device sampler then draws rank-independent slices [lo, hi) of ONE global ensemble
                seed64 = torch.randint(0, 2**62, (1,), dtype=torch.int64)
                if world > 1:
                    seed64 = seed64.to(device)
                    dist.broadcast(seed64, src=0)
                propagator.initial_conditions(q0, p0, Gamma_0, ntraj=hi - lo, ntraj_total=num_samples, index0=lo,
                                              seed=int(seed64.item()))
        _, err = guarded(install)
        agree(err)
        autocorrelation_ = np.zeros((nt,), dtype=complex)
        ic_correlation_ = np.zeros((nt,), dtype=complex)
        # t = 0 from the installed ensemble, then fused launches; a launch ends where the norm is due
        first = torch.tensor([[0.0, 0.0], [0.0, 0.0]], dtype=torch.float64, device=propagator.device)

        def first_values():
            c0 = complex(propagator.autocorrelation(energy0_es=en_zpt))
            k0 = complex(propagator.ic_correlation(potential, energy0_es=en_zpt))
            first[0, 0], first[0, 1], first[1, 0], first[1, 1] = c0.real, c0.imag, k0.real, k0.imag
        _, err = guarded(first_values)
        agree(err)
        if world > 1:
            dist.all_reduce(first)
        fh = first.cpu().numpy()
        autocorrelation_[0], ic_correlation_[0] = complex(fh[0, 0], fh[0, 1]), complex(fh[1, 0], fh[1, 1])
        t = 0
        while True:
            if calc_norm_every > 0 and t % calc_norm_every == 0:
                # all pairs of the GLOBAL ensemble (cli.py:424-429); sharded: ket vectors all-gathered, blocks all-reduced
                norm, err = guarded(lambda: propagator.norm(group=group))
                if isinstance(err, NotImplementedError):
                    if not norm_warned:
                        logger.warning(f"norm() is not available for this propagator ({err}); calc_norm_every is ignored")
                        norm_warned = True
                    err = None
                    if world > 1:
                        # the failing call skipped its collectives on every rank alike (the decision is rank independent)
                        pass
                agree(err)
                if norm is not None:
                    logger.info(f" time/fs= {times[t]*units.autime_to_fs}  norm= {norm:9.6f}")
            if t == nt - 1:
                break
            k = min(steps_per_launch, nt - 1 - t)
            if calc_norm_every > 0:
                k = min(k, calc_norm_every - t % calc_norm_every)

            def launch():
                a, i = propagator.propagate(potential, dt, k, energy0_es=en_zpt, group=group)
                assert not np.isnan(a).any(), f"encountered NaN's in autocorrelation : {a}"
                assert not np.isnan(i).any(), f"encountered NaN's in IC correlation : {i}"
                return a, i
            res, err = guarded(launch)
            agree(err)
            autocorrelation_[t + 1:t + 1 + k], ic_correlation_[t + 1:t + 1 + k] = res
            t += k
        # the last step() of the reference loop only advances the state, its correlations are never read (cli.py:436)

        def accumulate():
            data = dict(np.load(filename))
            ntraj_old, ntraj_new = data['trajectories'], num_samples
            ntraj_tot = ntraj_old + ntraj_new
            autocorrelation = (ntraj_new * autocorrelation_ + ntraj_old * data['autocorrelation']) / ntraj_tot
            ic_correlation = (ntraj_new * ic_correlation_ + ntraj_old * data['ic_correlation']) / ntraj_tot
            logger.info(f"<phi(0)|phi(0)>= {autocorrelation[0]}")
            assert abs(autocorrelation[0] - 1.0) < 1.0e-3
            data['trajectories'] = ntraj_tot
            data['autocorrelation'] = autocorrelation
            data['ic_correlation'] = ic_correlation
            data.pop('ic_rate', None)
            np.savez(filename, **data)
            return data
        data, err = guarded(accumulate) if rank == 0 else (None, None)
        agree(err)
    if world > 1:
        # every rank returns the arrays rank 0 wrote
        box = [data]
        dist.broadcast_object_list(box, src=0)
        data = box[0]
    return data
