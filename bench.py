#!/usr/bin/env python
"""bench.py -- FP64 trajectory-steps/s of the batched RK4 forward + discrete-adjoint sweep.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one single_shooting objective+gradient evaluation (functions/single_shooting.m:137-150:
compute_u -> compute_states -> compute_adjoints -> compute_dJdv) over the whole batch of
BASELINE.json configs[1]: 2^20 random initial conditions x 1000 RK4 steps of the bundled test
problem with a shared PWLinearControl (K = 101).  One trajectory-step = one RK4 step of one
trajectory, forward AND adjoint.  N > 1: every rank owns its own 2^20
trajectories (weak scaling) and J, x(T) are gathered to rank 0 over NCCL inside the timed region.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` goes through the
host-buffer C ABI with the copies inside the timed region; `roofline` is for the dominant
kernel; `cpu_baseline` is the CPU oracle (a C restatement of the MATLAB reference -- neither
MATLAB nor Octave exists in this image) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload: BASELINE.json configs[1] (SURVEY.md section 8d "cfg2")
BT = 1 << 20
N_STEPS = 1000
T_END = 10.0
K_PWL = 101
PARAMS = dict(c=1.5, m=3.0, r=0.05)          # tests/backprop_test.m:9-11
WORKLOAD = ("cfg2: batched single_shooting evaluation of tests/TestOCProblem.m, 2^20 random x0 x 1000 RK4 steps, "
            "shared PWLinearControl K=101, forward + discrete adjoint (J, dJdv)")

METRIC = "FP64 trajectory-steps/sec (fwd+adjoint RK4 sweep)"     # BASELINE.json's metric on configs[1]: one trajectory-step =
                                                                # one RK4 step of one trajectory, forward AND adjoint

# ---- algorithmic work per trajectory-step (add/mul/div/exp = 1, FMA = 2)
# TestOCProblem: C_F = 10, C_X = 8, C_U = 7, m = 2, nc = 1.  The control is shared by the batch, so
# its evaluation is per grid point, not per trajectory: c_u = 0.  c_v = 4 (PWLinear dJdv scatter).
FLOP_FWD = 4 * 10 + 13 * 2                    # 66
FLOP_REV = 4 * 8 + 14 * 2 + 4 * 7 + 2 * 1 + 2 * 4   # 98
BYTES_FWD = 8                                 # node state written once (n = 1)
BYTES_REV = 8                                 # ... and read once


# ---- other BASELINE.json configs, selectable with --config (parity cases; cfg2 is the bench line)
def make_workload(name, ocs, local_rank, bt_override=None):
    """Returns dict(integ, prob, ctl, n, K_rows, Bt, N_total, v (host, shared) or None, v_batch_shape,
    flop_fwd, flop_rev, bytes_fwd, bytes_rev, label, store_gb)."""
    if name == "cfg2":
        Bt = bt_override or BT
        tspan = ocs.matlab_linspace(0.0, T_END, N_STEPS + 1)
        integ = ocs.CudaRK4Integrator(tspan, device=local_rank)
        ctl = ocs.PWLinearControl(integ.t, K_PWL, 1)
        return dict(integ=integ, prob=ocs.TestOCProblem(PARAMS, [[0.0, 1.0]]), ctl=ctl, n=1, rows=K_PWL, Bt=Bt,
                    N_total=N_STEPS, v=np.random.default_rng(1).random(K_PWL), v_batch=None,
                    flop_fwd=FLOP_FWD, flop_rev=FLOP_REV, bytes_fwd=BYTES_FWD, bytes_rev=BYTES_REV, label=WORKLOAD,
                    kernels=("ocs_fwd_test_fast_u0_b0_p0", "ocs_rev_test_fast_u0_b1_p0"),
                    x0_lo=0.5, x0_hi=2.5)
    if name == "cfg2v":
        # cfg2 with one coefficient vector per trajectory (what a batched single_shooting solve evaluates):
        # control evaluated in the thread through the sliding two-knot window (c_u = 3 per grid point)
        Bt = bt_override or BT
        tspan = ocs.matlab_linspace(0.0, T_END, N_STEPS + 1)
        integ = ocs.CudaRK4Integrator(tspan, device=local_rank)
        ctl = ocs.PWLinearControl(integ.t, K_PWL, 1)
        return dict(integ=integ, prob=ocs.TestOCProblem(PARAMS, [[0.0, 1.0]]), ctl=ctl, n=1, rows=K_PWL, Bt=Bt,
                    N_total=N_STEPS, v=None, v_batch=(K_PWL, Bt), v_kind="uniform",
                    flop_fwd=FLOP_FWD + 2 * 3, flop_rev=FLOP_REV, bytes_fwd=BYTES_FWD, bytes_rev=BYTES_REV,
                    label="cfg2v: cfg2 with per-trajectory PWLinearControl coefficients v (101 x 2^20), forward + adjoint",
                    kernels=("ocs_fwd_test_fast_u1_b1_p0", "ocs_rev_test_fast_u1_b1_p0"), x0_lo=0.5, x0_hi=2.5)
    if name in ("cfg2p", "cfg2pv"):
        # parameter sweep: every trajectory its own (c, m, r) (tests/TestOCProblem.m:13-19) -- nothing is batch-uniform
        # any more, so exp(-r t) runs in the thread (4 per forward step, 8 per reverse step) and no table / weight
        # folding applies.  cfg2p: shared control coefficients; cfg2pv: per-trajectory coefficients as well.
        Bt = bt_override or BT
        tspan = ocs.matlab_linspace(0.0, T_END, N_STEPS + 1)
        integ = ocs.CudaRK4Integrator(tspan, device=local_rank)
        ctl = ocs.PWLinearControl(integ.t, K_PWL, 1)
        rng = np.random.default_rng(31)
        prob = ocs.TestOCProblem(dict(c=rng.uniform(1.3, 1.7, Bt), m=rng.uniform(2.8, 3.2, Bt), r=rng.uniform(0.03, 0.07, Bt)),
                                 [[0.0, 1.0]])
        pv = name == "cfg2pv"
        return dict(integ=integ, prob=prob, ctl=ctl, n=1, rows=K_PWL, Bt=Bt, N_total=N_STEPS,
                    v=None if pv else np.random.default_rng(1).random(K_PWL), v_batch=(K_PWL, Bt) if pv else None, v_kind="uniform",
                    flop_fwd=FLOP_FWD + (2 * 3 if pv else 0), flop_rev=FLOP_REV, bytes_fwd=BYTES_FWD + 24 / N_STEPS,
                    bytes_rev=BYTES_REV + 24 / N_STEPS,
                    label=name + ": parameter sweep of tests/TestOCProblem.m, 2^20 trajectories each with its own (c, m, r) and x0 "
                                 "x 1000 RK4 steps, " + ("per-trajectory" if pv else "shared") + " PWLinearControl K=101, forward + adjoint",
                    kernels=("ocs_fwd_test_fast_u%d_b%d_p1" % ((1, 1) if pv else (0, 0)),
                             "ocs_rev_test_fast_u%d_b1_p1" % (1 if pv else 0)), x0_lo=0.5, x0_hi=2.5)
    if name == "cfg3":
        # 6-state generated problem, one Chebyshev (degree 16) coefficient set per control candidate
        Bt = bt_override or 100_000
        K = 17
        tspan = ocs.matlab_linspace(0.0, T_END, N_STEPS + 1)
        integ = ocs.CudaRK4Integrator(tspan, device=local_rank)
        prob = ocs.harvest_chain_problem()
        ctl = ocs.ChebyshevControl(integ.t, K, 2)
        fl = prob.flops
        cu = cv = 2 * K * 2
        return dict(integ=integ, prob=prob, ctl=ctl, n=6, rows=2 * K, Bt=Bt, N_total=N_STEPS, v=None,
                    v_batch=(2 * K, Bt),
                    flop_fwd=4 * fl["C_F"] + 13 * 7 + 2 * cu, flop_rev=4 * fl["C_X"] + 14 * 7 + 4 * fl["C_U"] + 2 * 2 + 2 * cv,
                    # forward: node states (48 B) + the control grid kept for the reverse sweep (32 B) written; reverse sweep:
                    # both read again + dJdu written (32 B); the projection reads dJdu once (32 B), 2 K nc flops per grid point
                    bytes_fwd=48 + 32, bytes_rev=48 + 32 + 32, flop_proj=2 * cv, bytes_proj=32, default_bt=100_000,
                    label="cfg3: generated 6-state/6-costate harvest-chain problem (make_from_symbolic CUDA backend, NVRTC), "
                          "100k control candidates x 1000 RK4 steps, per-candidate ChebyshevControl degree 16, forward + adjoint",
                    kernels=("ocs_fwd_gen_u1_b2_p0", "ocs_rev_gen_u2_b0_p0"), x0_lo=1.0, x0_hi=2.5)
    if name == "cfg4":
        # RK4InfiniteIntegrator: main horizon + tail under uStar, stored state grids of both in HBM
        Bt = bt_override or 262_144
        N = 10_000
        tspan = ocs.matlab_linspace(0.0, 100.0, N + 1)
        tspanE = ocs.matlab_linspace(100.0, 200.0, N + 1)
        integ = ocs.CudaRK4InfiniteIntegrator(tspan, tspanE, 0.723368780097983, device=local_rank)
        ctl = ocs.PWLinearControl(integ.t, K_PWL, 1)
        return dict(integ=integ, prob=ocs.TestOCProblem(PARAMS, [[0.0, 1.0]]), ctl=ctl, n=1, rows=K_PWL, Bt=Bt,
                    N_total=2 * N, v=np.full(K_PWL, 0.723368780097983), v_batch=None,
                    flop_fwd=FLOP_FWD, flop_rev=FLOP_REV, bytes_fwd=BYTES_FWD, bytes_rev=BYTES_REV,
                    label="cfg4: RK4InfiniteIntegrator, 262144 trajectories x (10000 main + 10000 tail) RK4 steps, "
                          "shared PWLinearControl K=101, forward + adjoint, stored state grids in HBM",
                    kernels=("ocs_fwd_test_fast_u0_b0_p0", "ocs_rev_test_fast_u0_b1_p0 (+ u0_b0 tail)"),
                    x0_lo=0.5, x0_hi=2.5)
    raise SystemExit(f"unknown --config {name}")


def _measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _bind_to_gpu_numa_node(index):
    """Restricts this process to the CPUs NVML reports as local to GPU `index`; returns a function that restores
    the previous affinity.  No-op when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        old = os.sched_getaffinity(0)
        cpus &= old
        if not cpus:
            return lambda: None
        os.sched_setaffinity(0, cpus)
        return lambda: os.sched_setaffinity(0, old)
    except Exception:
        return lambda: None


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._dev = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
            return
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._dev, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._dev)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


class CpuOracle:
    """The C restatement oracle on the bench workload with everything that is not the sweep (grid, basis matrix,
    control coefficients, initial states) built ONCE, outside any timed region."""

    def __init__(self, ntraj_max):
        from oracle import oracle_c, rk4_oracle
        oracle_c.build()
        self.oc = oracle_c
        self.tspan = rk4_oracle.matlab_linspace(0.0, T_END, N_STEPS + 1)
        self.B = rk4_oracle.PWLinearControl(rk4_oracle.RK4Integrator(self.tspan).t, K_PWL, 1).B
        self.v = np.random.default_rng(1).random(K_PWL)
        self.x0 = np.random.default_rng(2).uniform(0.5, 2.5, (1, ntraj_max))
        # every host thread this process may run on -- set explicitly: torchrun exports OMP_NUM_THREADS=1, which
        # would silently time the reference on one core
        try:
            self.nthreads = len(os.sched_getaffinity(0))
        except AttributeError:
            self.nthreads = os.cpu_count() or 1

    def run(self, ntraj, mode=1):
        """One forward + adjoint evaluation of the first ntraj trajectories; returns seconds."""
        x0 = np.ascontiguousarray(self.x0[:, :ntraj])
        t0 = time.perf_counter()
        self.oc.evaluate(list(PARAMS.values()), self.tspan, self.B, x0, self.v, mode=mode, nthreads=self.nthreads)
        return time.perf_counter() - t0

    def threads_used(self):
        return self.oc.num_threads()


COLLECTIVE = "gather of J and x(T) to rank 0 (NCCL send/recv)"


def main_config(Bt, world):
    """The `config` object of the bench line; the reference arm prints the same keys (plus `sample`)."""
    return {"workload": WORKLOAD, "trajectories_per_gpu": Bt, "rk4_steps": N_STEPS,
            "l2": "inputs larger than L2 (%.1f GB node-state store per GPU)" % (8e-9 * N_STEPS * Bt),
            "collective": COLLECTIVE if world > 1 else "none"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference is MATLAB only and neither
    MATLAB nor Octave is installed, so this times the oracle port (C restatement of RK4Integrator.m /
    PWLinearControl.m / TestOCProblem.m, OpenMP across trajectories, every host thread).  Each step is a bounded
    sample of the workload (2^17 of the 2^20 trajectories, all 1000 RK4 steps, forward + adjoint); one full 2^20
    step is timed once on top of that and reported as `full_step`.  Grid, basis and inputs are built before the
    timed region."""
    if rank != 0:
        return
    ntraj = 1 << 17
    oracle = CpuOracle(BT)
    for _ in range(args.warmup):
        oracle.run(4096)
    dt = 0.0
    for _ in range(args.steps):
        dt += oracle.run(ntraj)
    cores = oracle.threads_used()
    value = args.steps * ntraj * N_STEPS / dt
    full = oracle.run(BT)
    sample = f"{ntraj} of the 2^20 trajectories x {N_STEPS} RK4 steps per step (fwd + adjoint), inputs prebuilt"
    emit({
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "trajectory-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": dict(main_config(BT, args.gpus), sample=sample),
        "full_step": {"trajectories": BT, "seconds": full, "value": BT * N_STEPS / full, "unit": "trajectory-steps/s"},
        "cpu_baseline": {"value": value, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_REAL_STDOUT = None


def run_rows(args, ocs, local_rank):
    """--config rows: the rows either side of the path, each timed through the host API
    (wall clock around the call, copies included) on one GPU.  Not a driver bench line; recorded in profiles/."""
    import torch
    out = {}
    prob = ocs.TestOCProblem(PARAMS, [[0.0, 1.0]])

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps, r

    # (4) batched compute_equilibrium: one steady state per parameter set
    rng = np.random.default_rng(21)
    ns = 1 << 20
    c, m, r = rng.uniform(1.3, 1.7, ns), rng.uniform(2.8, 3.2, ns), rng.uniform(0.03, 0.07, ns)
    sweep = ocs.TestOCProblem(dict(c=c, m=m, r=r), [[0.0, 1.0]])
    integ0 = ocs.CudaRK4Integrator(np.array([0.0, 1.0]), device=local_rank)
    dt, res = timed(lambda: ocs.compute_equilibrium(sweep, [2.7], [2.2], [.7], r=r, integrator=integ0))
    integ0.close()
    out["compute_equilibrium"] = {"parameter_sets": ns, "ms": 1e3 * dt, "solves_per_s": ns / dt,
                                  "converged": int((res[5] == 1).sum()), "max_resnorm": float(res[3].max())}
    from oracle import rk4_oracle
    t0 = time.perf_counter()
    nref = 50
    for b in range(nref):
        rk4_oracle.solve_equilibrium(rk4_oracle.TestOCProblem(c[b], m[b], r[b]), [2.7], [2.2], [.7], r[b])
    out["compute_equilibrium"]["cpu_oracle_solves_per_s"] = nref / (time.perf_counter() - t0)

    # (1) device-side batched single_shooting: tests/backprop_test.m's grid and control, 2^16 initial states
    Bt = 1 << 16
    tspan = ocs.matlab_linspace(0.0, 10.0, 501)
    integ = ocs.CudaRK4Integrator(tspan, device=local_rank)
    ctl = ocs.PWLinearControl(integ.t, 21, 1)
    integ.attach_control(ctl)
    x0 = np.random.default_rng(2).uniform(0.5, 2.5, (1, Bt))
    Lb, Ub = ctl.compute_nlp_bounds(prob.ControlBounds)
    dt, ss = timed(lambda: integ.single_shooting(prob, x0, ctl.compute_initial_v(0.5), Lb, Ub, max_iter=300, tol=1e-6), reps=1)
    out["single_shooting"] = {"trajectories": Bt, "rk4_steps": 500, "K": 21, "ms": 1e3 * dt, "solves_per_s": Bt / dt,
                              "converged": int(ss["converged"].sum()), "mean_iters": float(ss["iters"].mean()),
                              "trajectory_steps_per_s": float(ss["iters"].max()) * Bt * 500 / dt}

    # (3) soln.x / soln.lam / soln.u of those minimisers at nINTERP_PTS = 1001 query times (fb_sweep.m:22)
    tq = np.linspace(0.0, 10.0, 1001)
    nb = 1 << 14
    L = ocs._lib
    lib = L.load()
    bufs = [torch.empty((nb, tq.size), dtype=torch.float64).pin_memory() for _ in range(3)]   # MATLAB 1 x nq x nb each
    tqf = L.f64(tq)
    dt, _ = timed(lambda: L.check(lib.ocs_b200_eval_solution(integ._h, 0, 0, nb, tq.size, L.dp(tqf), L.dp(bufs[0].numpy()),
                                                             L.dp(bufs[1].numpy()), L.dp(bufs[2].numpy())), integ._h))
    dtp, sol = timed(lambda: integ.eval_solution(tq, 0, nb))
    assert np.array_equal(sol["x"][0].T, bufs[0].numpy()) and np.array_equal(sol["u"][0].T, bufs[2].numpy())
    out["eval_solution"] = {"trajectories": nb, "queries": tq.size, "ms": 1e3 * dt,
                            "interpolated_values_per_s": 3 * nb * tq.size / dt,
                            "d2h_bytes": 3 * 8 * nb * tq.size, "ms_with_pageable_numpy_outputs": 1e3 * dtp}
    integ.close()

    # (2) fb_sweep with the continuous costate (v1 adjointRHS): one full solve of 2^16 trajectories
    integ = ocs.CudaRK4Integrator(ocs.matlab_linspace(0.0, 10.0, 1001), device=local_rank)
    u0 = np.full((1, 2001), 0.5)
    for mode in ("discrete", "continuous"):
        dt, fb = timed(lambda: integ.fb_sweep(prob, x0, u0, nSWEEPS=60, costate=mode), reps=1)
        integ.set_profiling(True)          # per-kernel times from one more solve (the first launches are loaded by now)
        integ.kernel_stats(reset=True)
        integ.fb_sweep(prob, x0, u0, nSWEEPS=60, costate=mode)
        ks = integ.kernel_stats(reset=True)
        integ.set_profiling(False)
        out["fb_sweep_" + mode] = {"trajectories": Bt, "rk4_steps": 1000, "ms": 1e3 * dt, "solves_per_s": Bt / dt,
                                   "converged": int(fb["converged"].sum()), "max_sweeps": int(fb["sweeps"].max()),
                                   "kernel_ms_forward": ks["ms_forward"], "kernel_ms_costate": ks["ms_reverse"]}
    integ.close()
    emit({"metric": "rows beside the RK4 path (SURVEY.md 8f), host API wall clock on one B200", "rows": out,
          "n_gpus": 1, "data": "synthetic"})


def run_fbsweep(args, ocs, local_rank, rank, world):
    """--config fbsweep: the forward-backward sweep iteration (functions/fb_sweep.m:79-115 on the RK4 grid) on
    cfg2's batch: per sweep one forward sweep and one costate sweep with the control of EVERY trajectory on the
    G grid points in HBM (16 B/step read + 16 B/step written on top of the 16 B/step state stream), ControlChar
    update and convergence measure fused into the reverse kernel.  Tolerances 0 => exactly S sweeps per step.
    One sweep of one RK4 step of one trajectory = one trajectory-step."""
    import ctypes
    import torch
    gen = args.config == "cfg3fb"
    Bt, S = (100_000 if gen and args.bt == BT else args.bt), 4
    tspan = ocs.matlab_linspace(0.0, T_END, N_STEPS + 1)
    integ = ocs.CudaRK4Integrator(tspan, device=local_rank)
    kw = {}
    if gen:
        # BASELINE.json configs[2]: the generated 6-state problem, fb_sweep over 100k control candidates, each a
        # ChebyshevControl (degree 16) coefficient set; u0 = v*B is evaluated on the device
        prob = ocs.harvest_chain_problem()
        ctl = ocs.ChebyshevControl(integ.t, 17, 2)
        integ.attach_control(ctl)
        rng = np.random.default_rng(3 + rank)
        x0 = rng.uniform(1.0, 2.5, (6, Bt))
        v0 = 0.05 * rng.standard_normal((34, Bt))
        v0[:2] = rng.uniform(0.2, 0.8, (2, Bt))
        # host arrays as a MATLAB caller holds them: column-major (otherwise the Python mirror re-orders 27 MB per call)
        x0, v0 = np.asfortranarray(x0), np.asfortranarray(v0)
        u0, kw = None, dict(v0=v0)
    else:
        prob = ocs.TestOCProblem(PARAMS, [[0.0, 1.0]])
        x0 = np.random.default_rng(2 + rank).uniform(0.5, 2.5, (1, Bt))
        u0 = np.full((1, 2 * N_STEPS + 1), 0.5)
    nrep = max(1, min(args.steps, 5))
    # per-kernel device time: CUDA events around every launch, each kernel alone over the whole batch
    integ.set_profiling(True)
    for _ in range(max(1, min(args.warmup, 2))):
        integ.fb_sweep(prob, x0, u0, uRelTol=0.0, uAbsTol=0.0, nSWEEPS=S, **kw)
    integ.kernel_stats(reset=True)
    for _ in range(nrep):
        integ.fb_sweep(prob, x0, u0, uRelTol=0.0, uAbsTol=0.0, nSWEEPS=S, **kw)
    ks = integ.kernel_stats(reset=True)
    ms_f, ms_r = ks["ms_forward"] / (nrep * S), ks["ms_reverse"] / (nrep * S)
    # the timed region: the calls as a user makes them.  Device time of the sweep loops (events on the handle's stream
    # from the first kernel queued to the last one's end; a low-occupancy problem runs its sweeps as two lanes, so this is
    # less than the sum of the kernels timed alone) -> value; wall clock with the host copies -> e2e
    integ.set_profiling(False)
    integ.fb_sweep(prob, x0, u0, uRelTol=0.0, uAbsTol=0.0, nSWEEPS=S, **kw)
    integ.kernel_stats(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(nrep):
        out = integ.fb_sweep(prob, x0, u0, uRelTol=0.0, uAbsTol=0.0, nSWEEPS=S, **kw)
    dt = time.perf_counter() - t0
    ks = integ.kernel_stats(reset=True)
    assert (out["sweeps"] == S).all()
    ms_sweep = ks["ms_fb_loop"] / (nrep * S)          # one sweep = forward + costate/update kernels of the whole batch
    units = Bt * N_STEPS
    hbm_peak, hbm_src = _measured_peaks()
    L = ocs._lib.load()
    tf = ctypes.c_double()
    L.ocs_b200_measure_fp64_peak(local_rank, ctypes.byref(tf))
    # algorithmic per trajectory-step: forward 4*C_F + 13*m (control read from the grid) and 8 B state write +
    # 16 B control read; costate sweep 4*C_X + 14*m + 2 ControlChar (4 flop each) + change measure (2 x 5) and
    # 8 B state read + 16 B control read + 16 B control write
    ff, fr, bf, br = 4 * 10 + 13 * 2, 4 * 8 + 14 * 2 + 2 * 4 + 10, 24, 40
    tagk = "test_fast"
    if gen:     # m = 7, nc = 2: states 48 B/step, controls 32 B/step each way; ControlChar of two controls
        fl = prob.flops
        ff, fr, bf, br = 4 * fl["C_F"] + 13 * 7, 4 * fl["C_X"] + 14 * 7 + 2 * 2 * 4 + 2 * 10, 48 + 32, 48 + 32 + 32
        tagk = "gen"
    kf = {"kernel": "ocs_fwd_%s_u2_b0_p0" % tagk, "ms": ms_f, "tflops": units * ff / ms_f / 1e9, "gbs": units * bf / ms_f / 1e6}
    kr = {"kernel": "ocs_fbrev_%s_u2_b0_p0" % tagk, "ms": ms_r, "tflops": units * fr / ms_r / 1e9, "gbs": units * br / ms_r / 1e6}
    dom = kr if ms_r >= ms_f else kf
    fp_frac, hbm_frac = dom["tflops"] / tf.value, dom["gbs"] / hbm_peak
    bound = "hbm" if hbm_frac >= fp_frac else "fp64"
    emit({
        "metric": "FP64 trajectory-steps/sec (fb_sweep iteration on the RK4 grid)",
        "value": units * 1e3 / ms_sweep, "unit": "trajectory-steps/s", "n_gpus": 1, "steps": nrep, "warmup": 1,
        "ms_per_step": ms_sweep * S, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": ("cfg3fb: fb_sweep iteration of the generated 6-state/6-costate harvest-chain problem over 100k "
                                "control candidates (ChebyshevControl degree 16 each, evaluated on the device) x 1000 RK4 steps, "
                                "%d sweeps per step" % S) if gen else
                               "fbsweep: fb_sweep iteration of tests/TestOCProblem.m, 2^20 random x0 x 1000 RK4 steps, "
                               "per-trajectory control on the grid, %d sweeps per step" % S,
                   "trajectories_per_gpu": Bt, "rk4_steps": N_STEPS, "l2": "inputs larger than L2 (8.4 GB states + 16.8 GB controls)"},
        "gpu_launches": int(ks["launches"]),
        "roofline": {"bound": bound, "achieved": dom["gbs"] if bound == "hbm" else dom["tflops"],
                     "peak": hbm_peak if bound == "hbm" else tf.value, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                     "frac": max(hbm_frac, fp_frac), "kernel": dom["kernel"], "traffic": _committed_traffic("cfg3fb" if gen else "fbsweep", dom["kernel"]),
                     "peak_source": hbm_src,
                     "kernels": [kf, kr],
                     "step": {"ms_per_sweep": ms_sweep, "ms_kernels_alone": ms_f + ms_r,
                              "hbm_gbs": units * (bf + br) / ms_sweep / 1e6, "hbm_frac": units * (bf + br) / ms_sweep / 1e6 / hbm_peak,
                              "what": "one sweep (forward + costate/update) of the whole batch inside the solve loop: algorithmic "
                                      "bytes of both kernels over the loop's device time per sweep"}},
        "e2e": {"value": nrep * S * units / dt, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 8 * x0.size + 8 * (kw["v0"].size if gen else u0.size),
                "d2h_bytes_per_step": 8 * Bt * 2 + 4 * Bt * 2, "api": "ocs_b200_fb_sweep on host buffers"},
    })
    integ.close()


def run_solve(args, ocs, local_rank, rank, world):
    """--config cfg2solve: BASELINE.json configs[1] as literally stated -- the batched single_shooting SOLVE
    (functions/single_shooting.m:114) of 2^20 random initial conditions x 1000 RK4 steps, PWLinearControl K = 101,
    box bounds from compute_nlp_bounds -- through the host-buffer C ABI (ocs_b200_single_shooting): x0 goes up,
    vOpt, J, iters, converged come down, every objective/gradient evaluation stays on the device.  One bench "step"
    is one whole solve; a trajectory-step is one RK4 step (forward + adjoint) of one trajectory that was still
    open (converged trajectories are skipped by the sweeps and not counted)."""
    import torch
    import torch.distributed as dist
    Bt = args.bt
    dev = torch.device("cuda", local_rank)
    prob = ocs.TestOCProblem(PARAMS, [[0.0, 1.0]])
    integ = ocs.CudaRK4Integrator(ocs.matlab_linspace(0.0, T_END, N_STEPS + 1), device=local_rank)
    ctl = ocs.PWLinearControl(integ.t, K_PWL, 1)
    integ.attach_control(ctl)
    x0 = np.random.default_rng(2 + rank).uniform(0.5, 2.5, (1, Bt))
    Lb, Ub = ctl.compute_nlp_bounds(prob.ControlBounds)
    v0 = ctl.compute_initial_v(0.5)
    max_iter, tol = 200, 1e-6
    # pinned host buffers for the results (vOpt is 848 MB)
    restore = _bind_to_gpu_numa_node(local_rank)
    v_h = torch.empty((Bt, K_PWL), dtype=torch.float64).pin_memory()       # MATLAB (nc*K) x Bt, column-major
    J_h = torch.empty(Bt, dtype=torch.float64).pin_memory()
    it_h, cv_h = torch.empty(Bt, dtype=torch.int32).pin_memory(), torch.empty(Bt, dtype=torch.int32).pin_memory()
    v_h.view(-1)[::512].add_(0)
    restore()
    bufs = dict(v=v_h.numpy().T, J=J_h.numpy(), iters=it_h.numpy(), converged=cv_h.numpy())

    def solve():
        return integ.single_shooting(prob, x0, v0, Lb, Ub, max_iter=max_iter, tol=tol, out=bufs)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    for _ in range(max(1, min(args.warmup, 2))):
        ss = solve()
    reps = max(1, min(args.steps, 3))
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    integ.kernel_stats(reset=True)
    t0 = time.perf_counter()
    for _ in range(reps):
        ss = solve()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = integ.kernel_stats(reset=True)["launches"]
    # where a solve spends its device time: one more solve with CUDA events around every launch
    integ.set_profiling(True)
    ssp = solve()
    ksp = integ.kernel_stats(reset=True)
    integ.set_profiling(False)
    n_it = int(ssp["iters"].max())
    breakdown = {"iterations": n_it, "ms_forward": ksp["ms_forward"], "ms_reverse": ksp["ms_reverse"], "ms_update": ksp["ms_update"],
                 "ms_per_iteration": {"forward": ksp["ms_forward"] / (n_it + 1), "reverse": ksp["ms_reverse"] / n_it,
                                      "update": ksp["ms_update"] / n_it},
                 "update_gbs": 5 * 8.0 * K_PWL * Bt * n_it / (ksp["ms_update"] * 1e-3) / 1e9 if ksp["ms_update"] > 0 else None,
                 "what": "sums over one solve (the last forward sweep re-evaluates the minimisers); update = the projected-gradient "
                         "step, 5 passes over the (nc K) x Bt coefficient arrays",
                 # how many trajectories stopped after exactly k evaluations (rank 0's shard)
                 "evaluations_histogram": {str(k): int(c) for k, c in enumerate(np.bincount(ss["iters"].astype(np.int64))) if c}}
    evals = float(ss["iters"].astype(np.int64).sum())        # objective/gradient evaluations actually carried out
    t = torch.tensor([dt, evals, float(ss["converged"].sum()), float(ss["iters"].max())], dtype=torch.float64, device=dev)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, evals, nconv, itmax = float(tm[0]), float(t[1]), float(t[2]), float(tm[3])
    else:
        dt, evals, nconv, itmax = (float(z) for z in t)
    integ.close()
    if rank == 0:
        per = dt / reps
        val = evals * N_STEPS / per
        nv = K_PWL
        emit({"metric": "FP64 trajectory-steps/sec (batched single_shooting solve, fwd+adjoint RK4 sweeps)", "value": val,
              "unit": "trajectory-steps/s", "n_gpus": world, "steps": reps, "warmup": max(1, min(args.warmup, 2)),
              "ms_per_step": 1e3 * per, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
              "data": "synthetic",
              "config": {"workload": "cfg2solve: batched single_shooting solve of tests/TestOCProblem.m, 2^20 random x0 x 1000 "
                                     "RK4 steps, PWLinearControl K=101 with box bounds, spectral projected gradient on the "
                                     "device, tol 1e-6 on the projected gradient, host buffers in and out",
                         "trajectories_per_gpu": Bt, "rk4_steps": N_STEPS, "max_iter": max_iter,
                         "timing": "wall clock around the C-ABI call (copies included), max over ranks",
                         "l2": "inputs larger than L2 (%.1f GB node-state store per GPU)" % (8e-9 * N_STEPS * Bt)},
              "clocks": clocks, "gpu_launches": int(launches),
              "solve": {"solves_per_s": world * Bt / per, "converged": int(nconv), "trajectories": world * Bt,
                        "mean_evaluations": evals / (world * Bt), "max_evaluations": int(itmax), "device_time": breakdown},
              "e2e": {"value": val, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 8 * Bt + 8 * 3 * nv,
                      "d2h_bytes_per_step": 8 * Bt * (nv + 1) + 8 * Bt,
                      "api": "ocs_b200_single_shooting on host buffers: x0 up; vOpt, J, iters, converged down"}})
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def measure_fp64_peak(ocs, local_rank):
    """FP64 FMA peak by the DFMA micro-benchmark, with the SM clock sampled WHILE it runs and derived from the
    kernel's own cycle counter (cycles / elapsed): 148 SMs x 64 FMA/clk x 2 x clock is the nominal figure the
    measured one is compared with."""
    import ctypes
    L = ocs._lib.load()
    sampler = ClockSampler(local_rank)
    sampler.start()
    tf, mhz = ctypes.c_double(), ctypes.c_double()
    if hasattr(L, "ocs_b200_measure_fp64_peak_ex"):
        L.ocs_b200_measure_fp64_peak_ex(local_rank, ctypes.byref(tf), ctypes.byref(mhz))
    else:
        L.ocs_b200_measure_fp64_peak(local_rank, ctypes.byref(tf))
    clocks = sampler.stop()
    import torch
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    eff_mhz = mhz.value or (clocks["sm_mhz"] or 0)
    nominal = sms * 64 * 2 * eff_mhz * 1e6 / 1e12 if eff_mhz else None
    return tf.value, {"fp64_peak_tflops": tf.value, "fp64_peak_clock_mhz": mhz.value or None,
                      "fp64_peak_clock_mhz_nvml": clocks["sm_mhz"], "fp64_peak_reasons": clocks["reasons"],
                      "fp64_nominal_at_that_clock_tflops": nominal,
                      "fp64_peak_over_nominal": (tf.value / nominal) if nominal else None}


def pcie_ceiling(nbytes, dev, world):
    """What the host link gives this rank while EVERY rank downloads at once and nothing computes: one pinned
    device->host copy of `nbytes` per rank, started together.  The ceiling the end-to-end numbers are held against."""
    import torch
    import torch.distributed as dist
    n = nbytes // 8
    src = torch.empty(n, dtype=torch.float64, device=dev)
    dst = torch.empty(n, dtype=torch.float64).pin_memory()
    dst[::512] = 0
    best = 0.0
    for rep in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rep > 0:
            best = max(best, nbytes / dt / 1e9)
    del src, dst
    return best


def _committed_traffic(config, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this workload (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(config, {}).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def run_sub_config(name, extra=()):
    """One of the other BASELINE.json configurations in a subprocess of this script (its own CUDA context and
    memory); returns a compact record for the `configs` block of the main line."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--config", name, "--steps", "5", "--warmup", "3", "--no-cpu",
           "--no-configs", *extra]
    if name not in ("cfg3", "cfg3fb", "fbsweep", "cfg2solve"):      # (those carry a cheap end-to-end figure of their own)
        cmd.append("--no-e2e")
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        j = json.loads(line)
    except Exception as e:          # a failing side configuration must not take the bench line with it
        return {"error": repr(e)[:200]}
    rf = j.get("roofline", {})
    out = {"workload": j["config"]["workload"], "value": j["value"], "unit": j["unit"], "ms_per_step": j["ms_per_step"],
           "bound": rf.get("bound"), "frac": rf.get("frac"), "kernel": rf.get("kernel"),
           "kernels": [{k: q[k] for k in ("kernel", "ms", "tflops", "gbs") if k in q} for q in rf.get("kernels", [])],
           "wall_s": time.perf_counter() - t0}
    for k in ("step_fp64_frac", "step_hbm_frac", "step_ms_kernels_alone", "step"):
        if k in rf:
            out[k] = rf[k]
    if "solve" in j:
        out["solve"] = j["solve"]
    if "e2e" in j and j["e2e"]:
        out["e2e"] = j["e2e"].get("value")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bt", type=int, default=BT, help="trajectories per GPU (default 2^20)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the bounded runs of the other BASELINE configs")
    ap.add_argument("--gather", default="root", choices=["root", "peer", "all"],
                    help="N > 1: gather J, x(T) to rank 0 by NCCL send/recv (default), by peer copies over CUDA IPC + a 4-byte NCCL "
                         "completion all-reduce (measured slower at 8 GPUs: the all-reduce couples the ranks every step), or all-gather")
    ap.add_argument("--config", default="cfg2",
                    choices=["cfg2", "cfg2v", "cfg2p", "cfg2pv", "cfg2solve", "cfg3", "cfg3fb", "cfg4", "fbsweep", "rows"],
                    help="cfg2 (default) is BASELINE.json's bench configuration; cfg3/cfg4 are its other GPU configs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON): anything a library prints there (NCCL's version banner, ...)
    # is sent to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the RK4 sweeps have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    build = importlib.import_module("optimal-control-solvers_b200.build")
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    ocs = importlib.import_module("optimal-control-solvers_b200")
    from importlib import import_module
    distmod = import_module("optimal-control-solvers_b200.host.dist")

    if args.config in ("fbsweep", "cfg3fb"):
        run_fbsweep(args, ocs, local_rank, rank, world)
        return
    if args.config == "cfg2solve":
        run_solve(args, ocs, local_rank, rank, world)
        return
    if args.config == "rows":
        if rank == 0:
            run_rows(args, ocs, local_rank)
        return
    wl = make_workload(args.config, ocs, local_rank, None if args.bt == BT else args.bt)
    integ, prob, ctl, Bt, n_st, rows, NT = wl["integ"], wl["prob"], wl["ctl"], wl["Bt"], wl["n"], wl["rows"], wl["N_total"]
    integ.attach_control(ctl)
    stream = torch.cuda.Stream(device=dev)      # a real (non-default) stream: kernels, events and NCCL all use it
    torch.cuda.set_stream(stream)
    integ.set_stream(stream.cuda_stream)

    v = wl["v"]
    gen = torch.Generator(device=dev)
    gen.manual_seed(2 + rank)
    x0 = (wl["x0_lo"] + (wl["x0_hi"] - wl["x0_lo"]) *
          torch.rand((n_st, Bt), dtype=torch.float64, device=dev, generator=gen)).contiguous()
    v_dev = None
    if wl["v_batch"] is not None:     # per-trajectory coefficients: constant term U[0.2,0.8], higher ones N(0, 0.05^2)
        if wl.get("v_kind") == "uniform":
            v_dev = torch.rand(wl["v_batch"], dtype=torch.float64, device=dev, generator=gen)
        else:
            v_dev = 0.05 * torch.randn(wl["v_batch"], dtype=torch.float64, device=dev, generator=gen)
            v_dev[:2] = 0.2 + 0.6 * torch.rand((2, Bt), dtype=torch.float64, device=dev, generator=gen)
        v_dev = v_dev.contiguous()
    # J and x(T) live in one packed (1+n) x Bt buffer so that one gather moves both; double-buffered and gathered on
    # a side stream so that the transfer overlaps the next evaluation (host/dist.py).  Gather to rank 0 by default.
    pg = distmod.PipelinedGather(n_st, Bt, world, dev, rank=rank, mode=args.gather)
    status = torch.empty(Bt, dtype=torch.int32, device=dev)
    dJdv = torch.empty((rows, Bt), dtype=torch.float64, device=dev)
    lam0 = torch.empty((n_st + 1, Bt), dtype=torch.float64, device=dev)
    bound = [False]

    def evaluate_dev():
        res = pg.next_buffer()
        integ.evaluate_dev(prob, x0.data_ptr(), Bt, v=v, v_ptr=(v_dev.data_ptr() if v_dev is not None else 0),
                           J_ptr=res[0].data_ptr(), xT_ptr=res[1:].data_ptr(), status_ptr=status.data_ptr(),
                           dJdv_ptr=dJdv.data_ptr(), lam0_ptr=lam0.data_ptr(), bind=not bound[0])
        bound[0] = True
        return res

    def step():
        res = evaluate_dev()
        if world > 1:
            pg.gather(res)
        return res

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    integ.kernel_stats(reset=True)
    ev0.record(stream)
    for _ in range(args.steps):
        res = step()
    pg.finish()                 # every gather of the timed steps has completed before the closing event
    ev1.record(stream)
    sync_all()
    ms_local = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = integ.kernel_stats(reset=True)["launches"]    # sweeps + the table / fold kernels of every evaluation
    ms, per_rank, gather_check = ms_local, None, None
    if world > 1:
        t = torch.tensor([ms_local, float(clocks["sm_mhz"] or 0)], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per = [float(a[0]) for a in allt]
        ms = max(per)
        per_rank = {"ms_per_step_min": min(per) / args.steps, "ms_per_step_max": max(per) / args.steps,
                    "ms_per_step": [p / args.steps for p in per], "sm_mhz": [float(a[1]) for a in allt]}
        # what arrived at the root is what every rank computed: per-rank checksums of the last step's local J, x(T)
        # against checksums of the root's gathered slices (exact: same doubles, same summation order on one device type)
        loc = torch.stack([res[0].sum(), res[1:].sum(), res[0][Bt // 2], res[0].max()])
        alls = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(alls, loc)
        if rank == 0:
            out_root = pg.out[pg.k]
            ok = True
            for r in range(world):
                got = torch.stack([out_root[r, 0].sum(), out_root[r, 1:].sum(), out_root[r, 0][Bt // 2], out_root[r, 0].max()])
                ok = ok and bool(torch.equal(got, alls[r]))
            gather_check = {"gathered_equals_local": ok, "ranks": world, "bytes_per_rank_per_step": 8 * (1 + n_st) * Bt,
                            "mode": pg.mode}
            assert ok, "the J / x(T) gathered on rank 0 differ from what the ranks computed"
    value = world * Bt * NT * args.steps / (ms * 1e-3)

    # ---- per-kernel device time (CUDA events around every launch, on the launch stream)
    integ.set_profiling(True)
    nprof = max(3, min(args.steps, 10))
    for _ in range(nprof):
        evaluate_dev()
    torch.cuda.synchronize()
    ks = integ.kernel_stats(reset=True)
    integ.set_profiling(False)
    ms_fwd, ms_rev = ks["ms_forward"] / nprof, ks["ms_reverse_sweep"] / nprof
    ms_proj = ks["ms_project"] / nprof          # dense bases: dJdv = dJdu*B' is its own kernel after the reverse sweep

    out = None
    if rank == 0:
        tf_peak, peak_info = measure_fp64_peak(ocs, local_rank)
        hbm_peak, hbm_src = _measured_peaks()
        units = Bt * NT

        def kern(name, ms_k, flop, byt):
            return {"kernel": name, "ms": ms_k, "tflops": units * flop / (ms_k * 1e-3) / 1e12,
                    "gbs": units * byt / (ms_k * 1e-3) / 1e9}
        kf = kern(wl["kernels"][0], ms_fwd, wl["flop_fwd"], wl["bytes_fwd"])
        kr = kern(wl["kernels"][1], ms_rev, wl["flop_rev"] - wl.get("flop_proj", 0), wl["bytes_rev"])
        kernels = [kf, kr]
        if ms_proj > 0:
            kernels.append(kern("dense_project_kernel", ms_proj, wl.get("flop_proj", 0), wl.get("bytes_proj", 0)))
        dom = max(kernels, key=lambda k: k["ms"])
        fp_frac, hbm_frac = dom["tflops"] / tf_peak, dom["gbs"] / hbm_peak
        if fp_frac >= hbm_frac:
            roof = {"bound": "fp64", "achieved": dom["tflops"], "peak": tf_peak, "unit": "TFLOP/s", "frac": fp_frac,
                    "peak_source": "DFMA micro-benchmark in this run (ocs_b200_measure_fp64_peak_ex); no FP64 entry in MEASURED_PEAKS.json"}
        else:
            roof = {"bound": "hbm", "achieved": dom["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": hbm_frac,
                    "peak_source": hbm_src}
        traffic = None
        try:    # DRAM bytes per launch of the dominant kernel from the committed ncu capture (same workload only)
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if Bt == (wl.get("default_bt") or BT):
                traffic = tj.get(args.config, {}).get(dom["kernel"], {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
        roof.update({"kernel": dom["kernel"], "traffic": traffic,
                     "traffic_source": "profiles/traffic.json (ncu --set full capture of this workload, committed; not re-measured per run)",
                     "hbm": {"achieved": dom["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": hbm_frac, "peak_source": hbm_src},
                     "fp64": {"achieved": dom["tflops"], "peak": tf_peak, "unit": "TFLOP/s", "frac": fp_frac},
                     "kernels": kernels, **peak_info,
                     "algorithmic": {"flop_per_step_fwd": wl["flop_fwd"], "flop_per_step_rev": wl["flop_rev"],
                                     "flop_per_step_projection": wl.get("flop_proj", 0),
                                     "bytes_per_step_fwd": wl["bytes_fwd"], "bytes_per_step_rev": wl["bytes_rev"]}})
        # whole-evaluation view: (fwd + rev) algorithmic flops over the measured step time
        roof["step_fp64_frac"] = value / world * (wl["flop_fwd"] + wl["flop_rev"]) / 1e12 / tf_peak
        # ... and its algorithmic bytes over the same time.  The kernels above are timed alone over the whole batch; inside
        # the step a low-occupancy problem runs them as two overlapping lanes (ocs_b200.h, ocs_b200_evaluate_dev), so the
        # step is shorter than their sum
        step_bytes = wl["bytes_fwd"] + wl["bytes_rev"] + wl.get("bytes_proj", 0)
        roof["step_hbm_frac"] = value / world * step_bytes / 1e9 / hbm_peak
        roof["step_ms_kernels_alone"] = ms_fwd + ms_rev
        cfg = main_config(Bt, world) if args.config == "cfg2" else {
            "workload": wl["label"], "trajectories_per_gpu": Bt, "rk4_steps": NT,
            "l2": "inputs larger than L2 (%.1f GB node-state store per GPU)" % (8e-9 * n_st * NT * Bt),
            "collective": COLLECTIVE if world > 1 else "none"}
        out = {
            "metric": METRIC, "value": value,
            "unit": "trajectory-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "clocks": clocks, "gpu_launches": int(launches), "roofline": roof,
            # the two modes side by side, per GPU, from the per-kernel CUDA-event times
            "modes": {"fwd_only": {"value": units / (ms_fwd * 1e-3), "unit": "trajectory-steps/s", "ms": ms_fwd,
                                   "what": "compute_states alone (J, x(T))"},
                      "fwd_bwd": {"value": units / ((ms_fwd + ms_rev + ms_proj) * 1e-3), "unit": "trajectory-steps/s",
                                  "ms": ms_fwd + ms_rev + ms_proj, "what": "compute_states + compute_adjoints (J, dJdv)"}},
        }
        if per_rank:
            out["per_rank"] = per_rank
            out["gather_check"] = gather_check

    # ---- end to end through the host-buffer C ABI (rank-local; copies inside the timed region)
    e2e = None
    if not args.no_e2e and args.config in ("cfg2", "cfg4", "cfg3", "cfg2v", "cfg2p", "cfg2pv"):
        # pinned host buffers on the NUMA node of this rank's GPU (first touch under the GPU's CPU affinity):
        # with several ranks per box the downloads otherwise cross the socket interconnect
        restore = _bind_to_gpu_numa_node(local_rank)
        # host arrays in MATLAB layout (column-major, trajectory index last): a C-ordered (Bt, rows) array
        x0_h = torch.empty((Bt, n_st), dtype=torch.float64).pin_memory()
        x0_h.copy_(x0.cpu().T)
        J_h = torch.empty(Bt, dtype=torch.float64).pin_memory()
        xT_h = torch.empty((Bt, n_st), dtype=torch.float64).pin_memory()
        st_h = torch.empty(Bt, dtype=torch.int32).pin_memory()
        dJdv_h = torch.empty((Bt, rows), dtype=torch.float64).pin_memory()   # MATLAB (nc*K) x Bt, column-major
        lam0_h = torch.empty((Bt, n_st + 1), dtype=torch.float64).pin_memory()
        for buf in (x0_h, J_h, xT_h, dJdv_h, lam0_h):
            buf.view(-1)[::512].add_(0)          # touch every page while the affinity is in place
        # the same arrays as a MATLAB caller holds them: ordinary pageable memory (every mxArray is)
        pg_bufs = dict(x0=np.array(x0_h.numpy()), J=np.empty(Bt), xT=np.empty((Bt, n_st)), st=np.empty(Bt, dtype=np.int32),
                       dJdv=np.zeros((Bt, rows)), lam0=np.zeros((Bt, n_st + 1)))
        restore()
        lib = ocs._lib
        # coefficients: shared (one column) or one column per trajectory, uploaded with every evaluation
        if v_dev is not None:
            v_h = torch.empty((Bt, rows), dtype=torch.float64).pin_memory()
            v_h.copy_(v_dev.cpu().T)
            vf, v_cols = v_h.numpy(), Bt
        else:
            vf, v_cols = lib.f64(v), 1
        hnd = integ._h
        L = lib.load()
        pin_bufs = dict(x0=x0_h.numpy(), J=J_h.numpy(), xT=xT_h.numpy(), st=st_h.numpy(), dJdv=dJdv_h.numpy(), lam0=lam0_h.numpy())

        def e2e_step(b):
            lib.check(L.ocs_b200_set_control_v(hnd, lib.dp(vf), v_cols), hnd)
            lib.check(L.ocs_b200_evaluate(hnd, lib.dp(b["x0"]), Bt, None, 0, lib.dp(b["J"]), lib.dp(b["xT"]),
                                          lib.ip(b["st"]), lib.dp(b["dJdv"]), lib.dp(b["lam0"])), hnd)

        def timed_e2e(b):
            ne = max(2, min(args.steps, 5))
            e2e_step(b)
            sync_all()
            t0 = time.perf_counter()
            for _ in range(ne):
                e2e_step(b)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return ne, dt
        ne, dt = timed_e2e(pin_bufs)
        nep, dtp = timed_e2e(pg_bufs)
        same = bool(np.array_equal(pg_bufs["J"], pin_bufs["J"]) and np.array_equal(pg_bufs["dJdv"], pin_bufs["dJdv"]))
        d2h_bytes = 8 * Bt * (1 + n_st + rows + n_st + 1) + 4 * Bt
        ceiling = pcie_ceiling(d2h_bytes, dev, world)           # GB/s per rank with every rank copying at once
        if world > 1:
            t = torch.tensor([ceiling], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ceiling = float(t.item())
        per_rank_gbs = d2h_bytes * ne / dt / 1e9
        e2e = {"value": world * Bt * NT * ne / dt, "unit": "trajectory-steps/s",
               "h2d_bytes_per_step": 8 * Bt * n_st + 8 * rows * v_cols,
               "d2h_bytes_per_step": d2h_bytes,
               "steps": ne, "ms_per_step": 1e3 * dt / ne,
               "api": "ocs_b200_evaluate (nlpObjective: compute_states + compute_adjoints) on pinned host buffers; J, xT, status, dJdv, lam0 returned",
               # the D2H link with all ranks downloading and nothing computing, and how much of it the call reaches
               "pcie_ceiling_gbs": ceiling, "d2h_gbs_per_rank": per_rank_gbs,
               "frac_of_pcie_ceiling": per_rank_gbs / ceiling if ceiling else None,
               "e2e_pageable": {"value": world * Bt * NT * nep / dtp, "unit": "trajectory-steps/s", "ms_per_step": 1e3 * dtp / nep,
                                "vs_pinned": (dtp / nep) / (dt / ne), "results_identical_to_pinned": same,
                                "what": "the same call on ordinary pageable NumPy arrays (what MATLAB mxArrays are): "
                                        "transfers go through the handle's pinned ring and copy threads"}}

    # ---- the whole batched single_shooting solve end to end (configs[1] as literally stated): x0 up, vOpt down once
    solve = None
    if not args.no_e2e and args.config == "cfg2":
        ctl_s = ocs.PWLinearControl(integ.t, K_PWL, 1)
        Lb, Ub = ctl_s.compute_nlp_bounds(prob.ControlBounds)
        x0np = x0.cpu().numpy()
        v_out = torch.empty((Bt, K_PWL), dtype=torch.float64).pin_memory()
        sb = dict(v=v_out.numpy().T, J=np.empty(Bt), iters=np.empty(Bt, dtype=np.int32), converged=np.empty(Bt, dtype=np.int32))
        integ.single_shooting(prob, x0np, ctl_s.compute_initial_v(0.5), Lb, Ub, max_iter=200, tol=1e-6, out=sb)
        sync_all()
        dts_all = []
        for _ in range(3):          # one solve is ~0.2 s of wall clock: the median of three
            t0 = time.perf_counter()
            ss = integ.single_shooting(prob, x0np, ctl_s.compute_initial_v(0.5), Lb, Ub, max_iter=200, tol=1e-6, out=sb)
            torch.cuda.synchronize()
            dts_all.append(time.perf_counter() - t0)
        dts = sorted(dts_all)[1]
        ev_local = float(ss["iters"].astype(np.int64).sum())
        t = torch.tensor([dts, ev_local, float(ss["converged"].sum())], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dts, ev_all, nconv = float(tm[0]), float(t[1]), float(t[2])
        else:
            dts, ev_all, nconv = (float(z) for z in t)
        solve = {"value": ev_all * N_STEPS / dts, "unit": "trajectory-steps/s", "ms_per_solve_batch": 1e3 * dts,
                 "solves": world * Bt, "solves_per_s": world * Bt / dts, "converged": int(nconv),
                 "mean_evaluations": ev_all / (world * Bt),
                 "d2h_bytes": 8 * Bt * (K_PWL + 1) + 8 * Bt, "h2d_bytes": 8 * Bt,
                 "api": "ocs_b200_single_shooting on host buffers: x0 up; vOpt, J, iters, converged down once per solve "
                        "(every objective/gradient evaluation stays on the device); wall clock (median of 3 solves), max over ranks",
                 "ms_all": [1e3 * q for q in dts_all]}

    # ---- the same workload through ONE multi-device handle of the C ABI (rank 0 drives every GPU of the job)
    multi = None
    if world > 1 and args.config == "cfg2" and not args.no_e2e:
        integ.close()
        del dJdv, lam0, x0, pg
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 0:
            mi = ocs.CudaRK4Integrator(ocs.matlab_linspace(0.0, T_END, N_STEPS + 1), device=list(range(world)))
            mi.attach_control(ocs.PWLinearControl(mi.t, K_PWL, 1))
            x0m = np.random.default_rng(2).uniform(0.5, 2.5, (1, world * Bt))
            Jr = torch.empty(world * Bt, dtype=torch.float64, device="cuda:0")
            xTr = torch.empty((1, world * Bt), dtype=torch.float64, device="cuda:0")
            mi.evaluate_resident(prob, x0m, v, gradient=True)
            for _ in range(3):
                mi.evaluate_resident(prob, None, None, gradient=True)
                mi.gather_results_dev(Jr.data_ptr(), xTr.data_ptr())
            mi.synchronize()
            t0 = time.perf_counter()
            nm = max(5, min(args.steps, 20))
            for _ in range(nm):
                mi.evaluate_resident(prob, None, None, gradient=True)
                mi.gather_results_dev(Jr.data_ptr(), xTr.data_ptr())
            mi.synchronize()
            dtm = time.perf_counter() - t0
            ref = mi.results(gradient=False)
            okm = bool(np.array_equal(Jr.cpu().numpy(), ref["J"]) and np.array_equal(xTr.cpu().numpy(), ref["xT"]))
            multi = {"value": world * Bt * N_STEPS * nm / dtm, "unit": "trajectory-steps/s", "ms_per_step": 1e3 * dtm / nm,
                     "devices": world, "gathered_equals_host_results": okm,
                     "gather": os.environ.get("OCS_B200_GATHER", "nccl"),
                     "api": "ocs_b200_create_multi + ocs_b200_evaluate_resident + ocs_b200_gather_results_dev: one process, "
                            "one handle, a worker thread per GPU, J and x(T) gathered to the root device every step; wall clock"}
            mi.close()
        # the other ranks wait on the HOST (a key of the rendezvous store): an NCCL barrier would be a kernel spinning on
        # their GPU while rank 0's multi-device handle is sweeping on it
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("ocs_multi_done", "1")
        else:
            store.wait(["ocs_multi_done"])
        dist.barrier()

    if rank == 0:
        out["e2e"] = e2e
        if solve:
            out["solve"] = solve
        if multi:
            out["c_abi_multi"] = multi
        if not args.no_cpu and args.config == "cfg2":
            ntraj = 1 << 17
            oracle = CpuOracle(ntraj)
            oracle.run(4096)
            dtc = oracle.run(ntraj)
            out["cpu_baseline"] = {"value": ntraj * N_STEPS / dtc, "unit": "trajectory-steps/s", "cores": oracle.threads_used(), "kind": "port",
                                   "sample": f"{ntraj} of the 2^20 trajectories x {N_STEPS} steps, fwd + adjoint, inputs prebuilt, "
                                             "C restatement of the MATLAB reference (no MATLAB/Octave in the image), OpenMP"}
        if args.config == "cfg2" and world == 1 and not args.no_configs:
            # the other BASELINE.json configurations (and the solve), each a short run in its own process, so that the
            # driver's bench record carries them too.  Bounded: ~10-20 s each.
            try:
                integ.close()
            except Exception:
                pass
            torch.cuda.empty_cache()
            t_cfg = time.perf_counter()
            blocks = {}
            for name in ("cfg3", "cfg3fb", "cfg4", "fbsweep", "cfg2v", "cfg2p", "cfg2solve"):
                if time.perf_counter() - t_cfg > 200:
                    blocks[name] = {"skipped": "time budget of the configs block used up"}
                    continue
                blocks[name] = run_sub_config(name)
            out["configs"] = blocks
        emit(out)
    try:
        integ.close()
    except Exception:
        pass
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
