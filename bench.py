#!/usr/bin/env python
"""
bench.py -- headline benchmark of the HK hot path (BASELINE.json: "HK trajectory-steps/sec, AS 60-mode fp64").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--ntraj NTOTAL] [--dim D] [--dense]

Workload (configs[3]): synthetic anharmonic AS model, 60 modes, Herman-Kluk propagator,
10^6 trajectories in total, sharded contiguously over the N ranks (strong scaling: the global ensemble is fixed).
One "step" = one RK4 time step of every trajectory incl. 4 potential evaluations, the complex LU prefactor with
branch tracking and the contributions to both correlation functions.  The K timed steps run as ONE fused launch
(the state of a trajectory stays in shared memory for all K steps); the per-step correlation sums are
all-reduced over NCCL inside the timed region when N > 1.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port (oracle/sc_oracle.c, OpenMP over all
host cores) on a bounded sample of the same workload -- the reference itself is pure Python and does not exist on
the GPU box.
"""
import argparse
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

# DRAM traffic of the dominant kernel per trajectory-step, from the committed `ncu --set full` captures
# (dram__bytes_read.sum + dram__bytes_write.sum;
#  profiles/ncu_r01_f_k_rk4_wcols.txt: one launch of 9 768 trajectories x 10 steps)
TRAFFIC_BYTES_PER_TRAJ_STEP = {"k_rk4_wcols": (1528.7e6 + 7494.6e6) / 97680.0}
FLOP_PER_TRAJ_STEP = lambda d, dr, dense: 16.0 * d**3 + (8.0 / 3.0) * dr**3 + (8.0 * dr * d * d + 8.0 * dr * dr * d if dense else 0.0)  # noqa: E731


def fp64_peak_tflops():
    """FP64 roofline denominator: MEASURED_PEAKS.json has no fp64 entry, so the DMMA/DFMA microbenchmark of
    tools/fp64_peak.cu measured on this pool's B200 (profiles/fp64_peak_r01.json) is used"""
    path = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    try:
        with open(path) as f:
            j = json.load(f)
        return max(v for k, v in j.items() if k.startswith("dmma884_tflops")), "profiles/fp64_peak_r01.json (tools/fp64_peak.cu, DMMA.8x8x4 register-resident)"
    except Exception:
        return 37.0, "nominal B200 FP64 (fallback, microbenchmark file missing)"


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


def cpu_run(model, G, Q, q0, p0, ntraj, nsteps, nthreads=0, seed=0):
    """time the oracle port on `ntraj` trajectories x `nsteps` steps; returns (traj-steps/s, threads)"""
    from oracle import oracle
    if Q is None:
        pot = oracle.Potential.morse(model.omega, model.chi, model.nac)
    else:
        pot = oracle.Potential.rotated_morse(model.omega, model.chi, model.nac, Q)
    consts = oracle.Consts(G, G, G, q0, p0)
    zi, probi = oracle.sample_ensemble(G, G, q0, p0, ntraj, np.random.default_rng(seed))
    dt, _ = workloads.test_time_grid()
    t0 = time.perf_counter()
    oracle.run(pot, consts, zi, probi, dt, nsteps, model.en_zpt, nthreads=nthreads, want_state=False)
    el = time.perf_counter() - t0
    return ntraj * nsteps / el, oracle.lib().sc_oracle_num_threads(), el


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ntraj", type=int, default=1000000, help="global ensemble size")
    ap.add_argument("--dim", type=int, default=60)
    ap.add_argument("--dense", action="store_true", help="rotated AS model: dense Hessian and dense Gamma")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mma", action="store_true", help="force the DFMA kernel (diagnostics)")
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
              "sharding": f"{world} rank(s), contiguous trajectory slices",
              "l2": "state (>=116 KB/trajectory) far exceeds the 126 MB L2; no flush needed"}
    metric, unit = "HK trajectory-steps/sec, AS 60-mode fp64", "trajectory-steps/s"

    # ------------------------------------------------------------------ reference arm (CPU oracle port)
    if args.impl == "reference":
        if rank != 0:
            return
        n_s = 1024 if d >= 32 else 8192
        times = []
        for _ in range(W):
            cpu_run(model, G, Q, q0, p0, n_s, 1)
        t_all0 = time.perf_counter()
        for _ in range(K):
            v, cores, el = cpu_run(model, G, Q, q0, p0, n_s, 1)
            times.append(el)
        total = time.perf_counter() - t_all0
        value = n_s * K / sum(times)
        sample = f"{n_s} trajectories x 1 time step per bench step (bounded sample of the {args.ntraj}-trajectory workload)"
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
    from semiclassical_b200 import potentials, propagators

    T = lambda x: torch.from_numpy(np.ascontiguousarray(x))  # noqa: E731
    n_total = args.ntraj
    from semiclassical_b200 import distributed
    lo, hi = distributed.shard_bounds(n_total, rank, world)
    n_local = hi - lo
    if Q is None:
        pot = potentials.MorsePotential(T(model.omega), T(model.chi), T(model.nac))
    else:
        pot = potentials.RotatedMorsePotential(T(model.omega), T(model.chi), T(model.nac), T(Q))
    dt = workloads.test_time_grid()[0]
    if args.no_mma:
        os.environ["SC_NO_MMA"] = "1"
    pr = propagators.HermanKlukPropagator(T(G), T(G), device=device)
    # every rank samples its own shard with the propagator's sampler (initial_conditions); the shard is then kept in
    # pinned host memory so that the end-to-end leg starts from host buffers
    torch.manual_seed(1234 + rank)
    pr.initial_conditions(T(q0), T(p0), T(G), ntraj=n_local, ntraj_total=n_total)
    zi_pin, probi_pin = pr.zi.cpu().pin_memory(), pr.probi.cpu().pin_memory()
    ens_bytes = zi_pin.numel() * 8 + probi_pin.numel() * 8

    def install():
        pr.set_ensemble(T(q0), T(p0), T(G), zi_pin.to(device, non_blocking=True), probi_pin.to(device, non_blocking=True),
                        ntraj_total=n_total)

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
    auto, ic = run_steps(K)
    e1.record()
    barrier()
    tw1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = pr.launch_count() - l0
    value = n_total * K / (ms * 1e-3)
    # ---- timed region 2 (e2e): host buffers -> device, K steps, correlation functions back on the host
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    install()
    auto2, ic2 = run_steps(K)
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    clocks = sampler.stop(tw0, time.time()) if rank == 0 else None
    # ---- kernel-only duration of the dominant kernel, CUDA events on the launching stream
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from semiclassical_b200 import _native
    import ctypes
    st = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    handle = pot._handle(device)
    _native.check(_native.lib().sc_engine_set_timing(pr._engine, 1))
    e4.record()
    _native.check(_native.lib().sc_engine_step_dev(pr._engine, handle, dt, K, None, st))
    e5.record()
    torch.cuda.synchronize()
    ms_kernel = e4.elapsed_time(e5)
    kt = np.zeros(4)
    _native.check(_native.lib().sc_engine_get_timing(pr._engine, kt.ctypes.data))
    _native.check(_native.lib().sc_engine_set_timing(pr._engine, 0))
    pr.t = pr.t + K * dt
    flop = FLOP_PER_TRAJ_STEP(d, d, args.dense)
    peak, peak_src = fp64_peak_tflops()
    achieved_step = flop * n_local * K / (ms_kernel * 1e-3) / 1e12
    if kt[1] > 0.0:
        # column-chunked path: the dominant kernel is the RK4/monodromy kernel (16 d^3 of the 16 d^3 + 8/3 d^3 flops)
        dom_kernel, dom_ms, dom_flop = pr.kernel_name().split("+")[0], float(kt[1]), 16.0 * d**3
    else:
        dom_kernel, dom_ms, dom_flop = pr.kernel_name(), ms_kernel, flop
    achieved = dom_flop * n_local * K / (dom_ms * 1e-3) / 1e12
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    assert abs(c0 - 1.0) < 1e-3 * max(1.0, 3000.0 / np.sqrt(n_total)), f"C(0) = {c0}"
    assert np.all(np.isfinite(auto)) and np.all(np.isfinite(ic))
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": dict(config, steps_per_launch=K, kernel=pr.kernel_name(),
                                                                 trajectories_per_gpu=n_local),
            "e2e": {"value": n_total * K / (ms_e2e * 1e-3), "unit": unit,
                    "h2d_bytes_per_step": int(ens_bytes / K), "d2h_bytes_per_step": 40,
                    "note": "ensemble upload from pinned host memory + state initialisation + K steps + correlation functions to host"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_BYTES_PER_TRAJ_STEP.get(dom_kernel, None) and TRAFFIC_BYTES_PER_TRAJ_STEP[dom_kernel] * n_local * K,
                         "kernel": dom_kernel, "kernel_ms": dom_ms, "flop_per_trajectory_step": dom_flop, "peak_source": peak_src,
                         "whole_step": {"achieved": achieved_step, "frac": achieved_step / peak, "ms": ms_kernel,
                                        "flop_per_trajectory_step": flop, "kernels": pr.kernel_name(),
                                        "kernel_ms": {"qp_path": float(kt[0]), "rk4": float(kt[1]), "lu": float(kt[2]), "finish": float(kt[3])}},
                         "note": "FP64 pipe (DMMA/DFMA, measured 37.17 TFLOP/s); achieved = algorithmic flops of the dominant kernel "
                                 "(16 d^3 per trajectory-step: 4 RK4 stages x H [Mqq|Mqp]) / its device time from CUDA events on the "
                                 "launching stream; whole_step = (16 + 8/3) d^3 over ALL kernels of the step. Padding (60->64 rows) and "
                                 "structural zeros are not counted. traffic = DRAM bytes of the dominant kernel from the committed ncu "
                                 "capture (profiles/), scaled to this launch"},
            "check": {"C0": [c0.real, c0.imag], "auto_last": [auto[-1].real, auto[-1].imag]}}
    if not args.no_cpu_baseline:
        n_s = 1024 if d >= 32 else 8192
        ns = 10
        v, cores, el = cpu_run(model, G, Q, q0, p0, n_s, ns)
        line["cpu_baseline"] = {"value": v, "unit": unit, "cores": cores, "kind": "port",
                                "sample": f"{n_s} trajectories x {ns} steps of the same model ({el:.1f} s of CPU work), oracle/sc_oracle.c with OpenMP"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
