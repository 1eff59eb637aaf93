#!/usr/bin/env python
"""
bench.py -- headline benchmark of the HK hot path (BASELINE.json: "HK trajectory-steps/sec, AS 60-mode fp64").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--ntraj NTOTAL] [--dim D] [--dense]
                    [--dense-engine] [--no-dense-legs] [--no-cpu-baseline]

Workload (configs[3]): synthetic anharmonic AS model, 60 modes, Herman-Kluk propagator,
10^6 trajectories in total.  ONE global ensemble is drawn from a fixed seed (identically on every rank) and sharded
contiguously over the N ranks (strong scaling: the global ensemble is fixed), so `check` is the same at every N.
One "step" = one RK4 time step of every trajectory incl. 4 potential evaluations, the complex LU prefactor with branch
tracking and the contributions to both correlation functions.  The K timed steps run as ONE fused launch sequence; the
per-step correlation sums are all-reduced over NCCL inside the timed region when N > 1.

The JSON line carries
  roofline        the dominant kernel of the timed region (the structured pipeline of sc_chunk.cuh: H_s = H0 + diag(h_s)
                  with H0 = 0 for the separable AS model -- labelled as such, it is NOT the dense-engine figure)
  roofline_dense  (N = 1) the general dense pipeline (sc_stream.cuh) measured in the same run on
                    as_d60_dense_engine : the same AS ensemble, Hessians expanded to full per-trajectory matrices
                    harmonic_d60        : dense harmonic molecule-like model, d = 60, d' = 54, dense width matrices
                    rotated_as_d60      : rotated AS model, per-trajectory dense Hessians + dense width matrices
  other_configs   (N = 1, informational) trajectory-steps/s of the remaining BASELINE.json configs through the same public API,
                  state resident: C1 (AS 5 modes, HK), C2 (AS 5 modes, WM), C3 (methylium, harmonic), C5 (synthetic sGDML N = 17)
  peak            FP64 tensor-pipe peak measured in this run (sc_measure_fp64_peak: DMMA.8x8x4 chains on every SM)
`--impl reference` times the CPU oracle port (oracle/sc_oracle.c, OpenMP over all host cores) on a bounded sample of
the same workload -- the reference itself is pure Python and does not exist on the GPU box.
"""
import argparse
import ctypes
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from semiclassical_b200 import workloads  # noqa: E402


def flop_per_traj_step(d, dr, dense_gamma):
    """algorithmic flops: RK4 16 d^3, LU 8/3 d'^3, dense widths + 8 d' d^2 + 8 d'^2 d"""
    return 16.0 * d**3 + (8.0 / 3.0) * dr**3 + ((8.0 * dr * d * d + 8.0 * dr * dr * d) if dense_gamma else 0.0)


def load_traffic():
    """DRAM bytes per trajectory-step of the matrix kernels from the committed `ncu --set full` captures
    (dram__bytes_read.sum + dram__bytes_write.sum / trajectory-steps of the captured launch), profiles/traffic_r02.json"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r02.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(object):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 6] or [r for (_, r) in self.rows if len(r) >= 6]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}


def build_model(dim, dense):
    model = workloads.as_synthetic(dim)
    G = np.diag(model.omega)
    Q = None
    q0, p0 = model.q0, model.p0
    if dense:
        Q = workloads.random_orthogonal(dim, 11)
        G = Q @ G @ Q.T
        G = 0.5 * (G + G.T)
        q0, p0 = Q @ q0, Q @ p0
    return model, G, Q, q0, p0


def cpu_run(model, G, Q, q0, p0, ntraj, nsteps, seed=0):
    """time the oracle port on `ntraj` trajectories x `nsteps` steps with ALL host cores (explicit thread count: torchrun
    exports OMP_NUM_THREADS=1); returns (traj-steps/s, threads, seconds)"""
    from oracle import oracle
    if Q is None:
        pot = oracle.Potential.morse(model.omega, model.chi, model.nac)
    else:
        pot = oracle.Potential.rotated_morse(model.omega, model.chi, model.nac, Q)
    consts = oracle.Consts(G, G, G, q0, p0)
    zi, probi = oracle.sample_ensemble(G, G, q0, p0, ntraj, np.random.default_rng(seed))
    dt, _ = workloads.test_time_grid()
    ncpu = os.cpu_count() or 1
    t0 = time.perf_counter()
    oracle.run(pot, consts, zi, probi, dt, nsteps, model.en_zpt, nthreads=ncpu, want_state=False)
    el = time.perf_counter() - t0
    return ntraj * nsteps / el, oracle.lib().sc_oracle_num_threads(), el


def other_configs(device):
    """throughput of the remaining BASELINE.json configs (C1, C2, C3, C5: parity-test cases, not the headline) through the
    public API with the state resident, ensembles sampled on the device; informational block of the N = 1 line"""
    import torch
    from semiclassical_b200 import workloads, potentials, propagators
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x))  # noqa: E731
    out = {}

    def run(tag, pr, pot, q0, p0, G0, n, dt, K, e0, reps=2):
        try:
            torch.manual_seed(0)
            pr.initial_conditions(T(q0), T(p0), T(G0), ntraj=n)
            pr.propagate(pot, dt, K, e0)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                pr.propagate(pot, dt, K, e0)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            out[tag] = {"ntraj": n, "dim": pr.dim, "steps_per_launch": K, "trajectory_steps_per_s": n * K / ms * 1e3,
                        "kernel": pr.kernel_name()}
        except Exception as e:                                   # never let an informational leg break the bench line
            out[tag] = {"error": repr(e)[:200]}
        finally:
            del pr
            gc.collect()
            torch.cuda.empty_cache()

    dt5, _ = workloads.test_time_grid()
    m = workloads.as_5modes(0.02)
    G = np.diag(m.omega)
    pot = potentials.MorsePotential(T(m.omega), T(m.chi), T(m.nac))
    run("C1 AS 5 modes HK", propagators.HermanKlukPropagator(T(G), T(G), device=device), pot, m.q0, m.p0, G, 1000000, dt5, 50, m.en_zpt)
    run("C2 AS 5 modes WM alpha=beta=500", propagators.WaltonManolopoulosPropagator(T(G), T(G), 500, 500, device=device), pot,
        m.q0, m.p0, G, 200000, dt5, 20, m.en_zpt)
    try:
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", "hk_methylium.npz"), allow_pickle=False))
        potm = potentials.MolecularHarmonicPotential.from_arrays(g['pos0'], g['energy0'], g['grad0'], g['hess0'], g['masses'], g['nac'],
                                                                 float(g['origin']))
        run("C3 methylium harmonic HK (d'=6)", propagators.HermanKlukPropagator(T(g['Gamma_i']), T(g['Gamma_t']), device=device), potm,
            g['q0'], g['p0'], g['Gamma_0'], 100000, float(g['dt']), 100, float(g['energy0_es']))
    except Exception as e:
        out["C3 methylium harmonic HK (d'=6)"] = {"error": repr(e)[:200]}
    try:
        model, pos = workloads.gdml_synthetic()
        dg = len(pos)
        potg = potentials.MolecularGDMLPotential.from_arrays(model, np.full(dg, 12.0 * 1822.888486192), 1.0e-3 * np.ones(dg))
        Gg = np.diag(np.full(dg, 20.0))
        run("C5 synthetic sGDML N=17 HK", propagators.HermanKlukPropagator(T(Gg), T(Gg), device=device), potg, pos, np.zeros(dg), Gg,
            20000, 0.5, 5, 0.0)
    except Exception as e:
        out["C5 synthetic sGDML N=17 HK"] = {"error": repr(e)[:200]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ntraj", type=int, default=1000000, help="global ensemble size")
    ap.add_argument("--dim", type=int, default=60)
    ap.add_argument("--dense", action="store_true", help="rotated AS model: dense Hessian and dense Gamma")
    ap.add_argument("--dense-engine", action="store_true", help="AS model through the general dense pipeline (sc_stream.cuh)")
    ap.add_argument("--no-dense-legs", action="store_true", help="skip the roofline_dense legs")
    ap.add_argument("--dense-ntraj", type=int, default=148000, help="ensemble of the harmonic / rotated roofline_dense legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mma", action="store_true", help="force the DFMA kernel k_hk_generic (diagnostics)")
    args = ap.parse_args()

    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # NCCL would print its version banner on stdout next to the JSON line
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 0)
    d = args.dim
    model, G, Q, q0, p0 = build_model(d, args.dense)
    workload = f"synthetic anharmonic AS model, {d} modes, HK, {args.ntraj} trajectories" + (" (rotated: dense Hessian/Gamma)" if args.dense else "")
    config = {"workload": workload, "ntraj_global": args.ntraj, "dim": d, "dt_au": workloads.test_time_grid()[0],
              "potential": "rotated_morse" if args.dense else "morse", "gamma": "dense" if args.dense else "diag(omega)",
              "sharding": f"{world} rank(s), contiguous slices of ONE global ensemble (seed 1234)",
              "l2": "state (>=116 KB/trajectory) far exceeds the 126 MB L2; no flush needed"}
    metric, unit = "HK trajectory-steps/sec, AS 60-mode fp64", "trajectory-steps/s"

    # ------------------------------------------------------------------ reference arm (CPU oracle port)
    if args.impl == "reference":
        if rank != 0:
            return
        n_s = 1024 if d >= 32 else 8192
        ns = 10                                     # time steps per bench step: the t = 0 prefactor is amortised as in a real run
        times = []
        for _ in range(min(W, 1)):
            cpu_run(model, G, Q, q0, p0, n_s, ns)
        t_all0 = time.perf_counter()
        cores = 1
        for _ in range(K):
            v, cores, el = cpu_run(model, G, Q, q0, p0, n_s, ns)
            times.append(el)
        total = time.perf_counter() - t_all0
        value = n_s * ns * K / sum(times)
        sample = f"{n_s} trajectories x {ns} time steps per bench step (bounded sample of the {args.ntraj}-trajectory workload), {cores} OpenMP threads"
        line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * sum(times) / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": total}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    torch.set_default_dtype(torch.float64)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from semiclassical_b200 import _native, distributed, potentials, propagators

    T = lambda x: torch.from_numpy(np.ascontiguousarray(x))  # noqa: E731
    st = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    n_total = args.ntraj
    lo, hi = distributed.shard_bounds(n_total, rank, world)
    n_local = hi - lo
    dt = workloads.test_time_grid()[0]
    if args.no_mma:
        os.environ["SC_NO_MMA"] = "1"

    # FP64 roofline denominator, measured now on this device
    pk = np.zeros(2)
    _native.check(_native.lib().sc_measure_fp64_peak(pk.ctypes.data, 3, st))
    peak, peak_src = float(pk[0]), ("measured in this run: sc_measure_fp64_peak (DMMA.8x8x4 register-resident chains on every SM, best of 3; "
                                    f"DFMA chains: {pk[1]:.2f} TFLOP/s); MEASURED_PEAKS.json has no FP64 entry")

    def timed_leg(pr, pot, e_zpt, nsteps, dr, dense_gamma, label):
        """K fused steps with per-kernel CUDA-event timing on the launching stream -> roofline block"""
        handle = pot._handle(device)
        _native.check(_native.lib().sc_engine_set_timing(pr._engine, 1))
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        _native.check(_native.lib().sc_engine_step_dev(pr._engine, handle, dt, nsteps, None, st))
        e5.record()
        torch.cuda.synchronize()
        ms_all = e4.elapsed_time(e5)
        kt = np.zeros(8)
        _native.check(_native.lib().sc_engine_get_timing_slots(pr._engine, kt.ctypes.data, 8))
        _native.check(_native.lib().sc_engine_set_timing(pr._engine, 0))
        pr.t = pr.t + nsteps * dt
        dd, n = pr.dim, pr.ntraj
        kname = pr.kernel_name()
        flop = flop_per_traj_step(dd, dr, dense_gamma)
        ach_step = flop * n * nsteps / (ms_all * 1e-3) / 1e12
        # dominant kernel: the RK4 / monodromy kernel; k_rk4_stream also carries the left factors of a dense prefactor
        dom_flop = 16.0 * dd**3 + (8.0 * dr * dd * dd if (dense_gamma and kname.startswith("k_rk4_stream")) else 0.0)
        dom_ms = float(kt[1]) if kt[1] > 0.0 else ms_all
        if kt[1] <= 0.0:
            dom_flop = flop
        ach = dom_flop * n * nsteps / (dom_ms * 1e-3) / 1e12
        tr = load_traffic().get(kname.split("+")[0])
        return {"label": label, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": (tr["bytes_per_trajectory_step"] * n * nsteps) if tr else None,
                "traffic_source": tr.get("capture") if tr else None,
                "kernel": kname.split("+")[0], "kernel_ms": dom_ms, "flop_per_trajectory_step": dom_flop,
                "trajectories": n, "steps": nsteps, "traj_steps_per_s": n * nsteps / (ms_all * 1e-3),
                "whole_step": {"achieved": ach_step, "frac": ach_step / peak, "ms": ms_all, "flop_per_trajectory_step": flop,
                               "kernels": kname,
                               "kernel_ms": {"path_and_overlap_terms": float(kt[0]), "potential_hessians": float(kt[5]), "rk4": float(kt[1]),
                                             "rmult": float(kt[4]), "lu": float(kt[2]), "finish": float(kt[3])}}}

    if Q is None:
        pot = potentials.MorsePotential(T(model.omega), T(model.chi), T(model.nac))
    else:
        pot = potentials.RotatedMorsePotential(T(model.omega), T(model.chi), T(model.nac), T(Q))
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=device)
    # ONE global ensemble from a fixed seed: the device sampler is counter based (Philox), trajectory i of the global ensemble is
    # a pure function of (seed, i), so every rank draws exactly its contiguous slice [lo, hi); the slice is kept in pinned host
    # memory so that the end-to-end leg starts from host buffers
    zi_s, probi_s = pr.sample_ensemble(T(q0), T(p0), T(G), n_local, index0=lo, seed=1234)
    zi_pin, probi_pin = zi_s.cpu().pin_memory(), probi_s.cpu().pin_memory()
    del zi_s, probi_s
    torch.cuda.empty_cache()
    ens_bytes = zi_pin.numel() * 8 + probi_pin.numel() * 8

    def install():
        pr.set_ensemble(T(q0), T(p0), T(G), zi_pin.to(device, non_blocking=True), probi_pin.to(device, non_blocking=True),
                        ntraj_total=n_total)

    install()
    if args.dense_engine:
        pr.set_option("dense_engine", 1)

    def run_steps(nsteps):
        # K fused steps; for N > 1 the (K, 5) buffer of per-step sums is all-reduced on the device (NCCL) before
        # the single device->host copy
        return pr.propagate(pot, dt, nsteps, model.en_zpt, group=True if dist is not None else None)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    c0 = pr.autocorrelation(model.en_zpt)
    if dist is not None:   # every rank holds its shard's share of C(0)
        t0 = torch.tensor([c0.real, c0.imag], device=device)
        dist.all_reduce(t0)
        c0 = complex(float(t0[0]), float(t0[1]))
    # warm-up: at least W time steps, in one launch of the same shape as the timed one (same scratch partition, same
    # kernels' step counts) so that nothing is allocated or configured inside the timed region
    if W > 0:
        run_steps(max(W, K))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- timed region 1: K fused steps, state resident in HBM
    barrier()
    l0 = pr.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record()
    run_steps(K)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = pr.launch_count() - l0
    value = n_total * K / (ms * 1e-3)
    # ---- timed region 2 (e2e): host buffers -> device, K steps, correlation functions back on the host.  It restarts from
    # the ensemble at t = 0, so its correlation functions are a property of the global ensemble alone (`check`)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    install()
    auto2, ic2 = run_steps(K)
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    clocks = sampler.stop(tw0, time.time()) if rank == 0 else None
    # ---- kernel-only durations, CUDA events on the launching stream
    barrier()
    dense_gamma = args.dense
    if args.dense or args.dense_engine:
        label = "general dense pipeline (sc_stream.cuh)"
    else:
        label = ("structured pipeline (sc_chunk.cuh): H_s = H0 + diag(h_s), dense base H0 multiplied in full on DMMA but identically "
                 "ZERO for the separable AS model -- the rate of the instruction stream, NOT a dense-engine figure (see roofline_dense)")
    roof = timed_leg(pr, pot, model.en_zpt, K, d, dense_gamma, label)
    roofline_dense = None
    if world == 1 and not args.no_dense_legs and not args.dense and d >= 17:
        roofline_dense = {}
        if not args.dense_engine:
            # (i) the same ensemble, same model, through the general dense engine
            pr.set_option("dense_engine", 1)
            run_steps(K)
            roofline_dense["as_d%d_dense_engine" % d] = timed_leg(
                pr, pot, model.en_zpt, K, d, False,
                "configs[3] ensemble on the general dense pipeline: per-trajectory stage Hessians expanded to full d x d matrices and "
                "streamed; the kernel does not know they are diagonal")
            pr.set_option("dense_engine", 0)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    if roofline_dense is not None:
        del pr
        gc.collect()
        torch.cuda.empty_cache()
        nd = args.dense_ntraj
        # (ii) dense harmonic molecule-like model: dense constant Hessian, dense rank-deficient width matrices (d' = d - 6)
        hm = workloads.harmonic_molecule_synthetic(d)
        hpot = potentials.MolecularHarmonicPotential.from_arrays(hm['pos0'], hm['energy0'], hm['grad0'], hm['hess0'], hm['masses'], hm['nac'])
        hpr = propagators.HermanKlukPropagator(T(hm['Gamma_0']), T(hm['Gamma_0']), device=device)
        torch.manual_seed(4321)
        hpr.initial_conditions(T(hm['q0']), T(hm['p0']), T(hm['Gamma_0']), ntraj=nd)
        hpr.propagate(hpot, dt, K, hm['en_zpt'])
        roofline_dense["harmonic_d%d" % d] = timed_leg(
            hpr, hpot, hm['en_zpt'], K, d - 6, True,
            "workloads.harmonic_molecule_synthetic: dense Hessian (constant, streamed from one copy), dense width matrices with 6 zero modes")
        del hpr
        gc.collect()
        torch.cuda.empty_cache()
        # (iii) rotated AS model: per-trajectory dense Hessians Q diag(h) Q^T and dense width matrices
        _, Gr, Qr, q0r, p0r = build_model(d, True)
        rpot = potentials.RotatedMorsePotential(T(model.omega), T(model.chi), T(model.nac), T(Qr))
        rpr = propagators.HermanKlukPropagator(T(Gr), T(Gr), device=device)
        torch.manual_seed(4322)
        rpr.initial_conditions(T(q0r), T(p0r), T(Gr), ntraj=nd)
        rpr.propagate(rpot, dt, K, model.en_zpt)
        roofline_dense["rotated_as_d%d" % d] = timed_leg(
            rpr, rpot, model.en_zpt, K, d, True,
            "rotated AS model (SURVEY 8c-vi): per-trajectory dense Hessians (formed by k_expand_hessian, not counted as algorithmic "
            "flops) and dense width matrices")
        del rpr
        gc.collect()
        torch.cuda.empty_cache()
    assert abs(c0 - 1.0) < 1e-3 * max(1.0, 3000.0 / np.sqrt(n_total)), f"C(0) = {c0}"
    assert np.all(np.isfinite(auto2)) and np.all(np.isfinite(ic2))
    # multi-GPU parity evidence: the e2e leg's correlation functions depend only on the global ensemble (seed 1234), so they
    # must agree at every N to reduction-order accuracy; profiles/check_r02.json holds the N = 1 values
    check = {"C0": [c0.real, c0.imag], "auto_last": [auto2[-1].real, auto2[-1].imag], "ic_last": [ic2[-1].real, ic2[-1].imag],
             "auto_abs_sum": float(np.abs(auto2).sum())}
    try:
        with open(os.path.join(ROOT, "profiles", "check_r02.json")) as f:
            stored = json.load(f).get(f"{'dense' if args.dense else 'as'}_d{d}_n{n_total}_k{K}")
        if stored:
            ref = complex(*stored["auto_last"])
            check["rel_diff_vs_stored_n1"] = abs(auto2[-1] - ref) / abs(ref)
            check["matches_stored_n1"] = bool(check["rel_diff_vs_stored_n1"] < 1e-9)
    except Exception:
        pass
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": dict(config, steps_per_launch=K, kernel=roof["whole_step"]["kernels"],
                                                                 trajectories_per_gpu=n_local),
            "e2e": {"value": n_total * K / (ms_e2e * 1e-3), "unit": unit,
                    "h2d_bytes_per_step": int(ens_bytes / K), "d2h_bytes_per_step": 40,
                    "note": "ensemble upload from pinned host memory + state initialisation + K steps + correlation functions to host"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": dict(roof, peak_source=peak_src,
                             note="achieved = algorithmic flops of the dominant kernel (16 d^3 per trajectory-step: 4 RK4 stages x "
                                  "H [Mqq|Mqp]; + 8 d' d^2 when k_rk4_stream also applies the left prefactor factors) / its device time from "
                                  "CUDA events on the launching stream; whole_step = all algorithmic flops (SURVEY 8d) over ALL kernels "
                                  "of the step. Padding (60->64 rows), Hessian formation and structural zeros are not counted. traffic = "
                                  "DRAM bytes per trajectory-step of the dominant kernel from the committed ncu capture, scaled to this launch"),
            "check": check}
    if roofline_dense is not None:
        line["roofline_dense"] = roofline_dense
        try:
            line["other_configs"] = other_configs(device)
        except Exception as e:                                   # informational block: never break the bench line
            line["other_configs"] = {"error": repr(e)[:200]}
    if not args.no_cpu_baseline and world == 1:
        n_s = 4096 if d >= 32 else 32768             # 10-30 s of CPU work on the box's host cores
        ns = 10
        v, cores, el = cpu_run(model, G, Q, q0, p0, n_s, ns)
        line["cpu_baseline"] = {"value": v, "unit": unit, "cores": cores, "kind": "port",
                                "sample": f"{n_s} trajectories x {ns} steps of the same model ({el:.1f} s of CPU work), oracle/sc_oracle.c with OpenMP"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
