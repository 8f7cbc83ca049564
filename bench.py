#!/usr/bin/env python
"""bench.py -- the contract benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode M]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): MPC QPs solved/sec at horizon N=16, batch 64K per GPU; p50 batch latency.
Workload: BASELINE configs[2] -- 65 536 randomized Go2 trot states + references per GPU, cold start
(every step re-solves the batch from scratch; inputs of one batch, 213 MB, exceed the 126 MB L2).
One "step" = one batched ``solve_QP`` (contact table + dynamics + QP data + exact active-set solve, fused on the device).

  value      QPs/s with inputs resident in HBM (CUDA events, barrier + sync on both sides, max over ranks)
  e2e        the same through the host-buffer C-ABI call ``cmpc_solve_host`` (pinned host memory in,
             forces/status out; H2D and D2H inside the timed region)
  e2e_cycle  the whole cycle from state + command through ``cmpc_cycle_host`` (trajectory generated on the device,
             first-step forces out: 408 B in / 96 B out per robot)
  roofline   the dominant kernel vs the FP64-FMA roofline at the flops of the route every robot took, its duration
             measured live by CUDA events inside ``cmpc_solve`` (peak measured on this GPU by the library's DFMA
             micro-benchmark -- MEASURED_PEAKS.json has no FP64 entry); ``kernels`` lists every kernel of the solve,
             ``hbm`` the achieved HBM GB/s vs MEASURED_PEAKS.json, ``traffic`` the ncu DRAM bytes of the newest capture
  cpu_baseline   the oracle's C port of the reference CPU path (OSQP restatement) on the host cores, with the
             CPU-1 / CPU-P warm variants of BASELINE.md section 2, and a sampled L2 check of the GPU forces

``--impl reference`` times the reference's own CasADi -> OSQP path when the casadi wheel and the reference tree are
present (oracle/live_reference.py), else that CPU port alone.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MPC QPs solved/sec at N=16, batch 64K per GPU (cold start, build+solve fused)"
UNIT = "QPs/s"
WORKLOAD = "BASELINE configs[2]: 65536 randomized Go2 states+references per GPU, trot 3 Hz duty 0.6, horizon 16"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500, help="timed steps (default 500: about 2 s of device time per arm)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="robots per GPU")
    ap.add_argument("--mode", default="active_set", choices=["active_set", "admm"])
    ap.add_argument("--stress", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--prepass", type=int, default=4, help="kernel ahead of the condensed kernel: 0 off, 1-3 Riccati sweeps of round 1 (nominal robots only), "
                                                           "4 wrench-space projected-Riccati active set for all robots (default)")
    ap.add_argument("--sweep", action="store_true", help="batch-size sweep 1..262144 -> gpurun_out/sweep.json")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons of one GPU while the timed region runs: NVML polled from a thread every 5 ms
    (nvidia_ml_py), else ``nvidia-smi -lms 20``."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.samples, self.stop_flag, self.nvml = [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        R = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
             "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
             "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
             "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = get_reasons(self.h)
                self.samples.append((float(sm), float(mx), pw, {k for k, bit in R.items() if rs & bit}))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nvml:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        sm, mx, pw, reasons = [], [], [], set()
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1)
            for a, b, c, r in self.samples:
                sm.append(a); mx.append(b); pw.append(c); reasons |= r
            src = "nvml 5 ms"
        else:
            if not self.proc:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            for ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            src = "nvidia-smi -lms 20"
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s_ for s_, p_ in zip(sm, pw) if p_ >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "source": src, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def l2_check(rec, u_gpu, n=16):
    """a QP counts as solved when its status says so AND a sampled subset passes L2 -- forces within
    1e-3 relative / 1e-2 N of the exact optimum of the reference's QP (oracle/exact.py; the oracle is the checker here,
    never the thing measured)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import force_error, oracle_solution
    idx = np.unique(np.linspace(0, rec.B - 1, n).astype(int))
    worst_abs, worst_rel = 0.0, 0.0
    for b in idx:
        o = oracle_solution(rec, int(b))
        err, rel = force_error(np.asarray(u_gpu[b]).reshape(-1), o["sol"]["U"])
        worst_abs, worst_rel = max(worst_abs, float(err)), max(worst_rel, float(rel))
    return {"n": int(len(idx)), "max_abs_err_N": worst_abs, "max_err_over_tolerance": worst_rel, "pass": bool(worst_rel < 1.0),
            "tolerance": "1e-2 N + 1e-3 |u*| per force component, against the oracle's exact optimum"}


def cpu_baseline(rec, budget_s=15.0, eps=1e-5):
    """Oracle C port (OSQP restatement on the reference's sparse QP) on the host cores, bounded sample."""
    from oracle import cpu_port
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    opts = cpu_port.default_opts(eps_abs=eps, eps_rel=eps)
    probe = rec.slice(0, min(16, rec.B))
    t = time.perf_counter()
    r1 = cpu_port.solve_batch(probe, opts, nthreads=1)
    t1 = time.perf_counter() - t
    rate1 = probe.B / t1
    n = int(min(rec.B, max(cores, rate1 * cores * budget_s)))
    sample = rec.slice(0, n)
    t = time.perf_counter()
    r = cpu_port.solve_batch(sample, opts, nthreads=cores)
    tp = time.perf_counter() - t
    variants = {}
    try:
        variants = cpu_variants(records_mod=None, rec=rec, opts=opts, cores=cores)
    except Exception as e:        # the variants never break the contract line
        variants = {"error": repr(e)}
    return {"value": n / tp, "unit": UNIT, "cores": cores, "kind": "port", "variants": variants,
            "value_1thread": rate1,
            "iters_mean": float(r["iters"].mean()), "solved_frac": float((r["status"] == 1).mean()),
            "sample": f"first {n} QPs of the same workload, cold start, OSQP restatement (oracle/osqp_port.c) on the "
                      f"reference's sparse 384-var QP, reference OPTS but eps_abs=eps_rel={eps:g}, {cores} threads "
                      f"({tp:.1f} s); single-thread probe {probe.B} QPs ({t1:.1f} s). Restated CPU baseline, not the "
                      f"reference binary (CasADi/OSQP not installable)"}


def cpu_variants(records_mod, rec, opts, cores):
    """The other CPU runs BASELINE.md section 2 lists, on the restated port (bounded: about 10 s in all):
    CPU-1  BASELINE configs[0]: one robot, sequential warm-started cycles at the reference's cadence (MPC every 20 ms,
           command schedule of test_MPC.py:37-47; the robot advanced between cycles by the SRB stand-in for MuJoCo,
           records.next_cycle), one thread: ms per QP p50 / p90 / max;
    CPU-P  warm: a 1 024-QP slice of the workload re-solved from the solution of a 20 ms-earlier cycle, all threads."""
    from convex_mpc_b200 import records
    from oracle import cpu_port
    out = {}
    one = records.random_records(1, seed=1, stress=0.0)
    sched = [(0.0, 0.7, 0.0, 0.0), (2.0, 0.0, 0.3, 0.0), (4.0, 0.0, 0.0, 2.0), (6.0, 0.6, 0.0, 2.0), (8.0, 0.8, 0.0, 0.0)]   # t, vx, vy, wz
    state, ms, its, ok = None, [], [], 0
    t_end = time.perf_counter() + 6.0
    cycles = 0
    for c in range(500):
        tnow = c * 0.02
        vx, vy, wz = [s_[1:] for s_ in sched if s_[0] <= tnow][-1]
        one = records.retarget(one, vx, vy, wz) if hasattr(records, "retarget") else one
        t = time.perf_counter()
        r = cpu_port.solve_batch(one, opts, warm=state is not None, state=state, nthreads=1)
        ms.append((time.perf_counter() - t) * 1e3)
        its.append(int(r["iters"][0])); ok += int(r["status"][0] == 1)
        state = (r["w"], r["y"], r["rho"])
        N = one.N
        one = records.next_cycle(one, r["w"][:, 12 * N:12 * N + 12])
        cycles += 1
        if time.perf_counter() > t_end:
            break
    out["cpu1_sequential_warm"] = {"cycles": cycles, "ms_p50": float(np.percentile(ms, 50)), "ms_p90": float(np.percentile(ms, 90)),
                                   "ms_max": float(np.max(ms)), "qps": cycles / (np.sum(ms) * 1e-3), "iters_mean": float(np.mean(its)),
                                   "solved": ok, "threads": 1,
                                   "what": "one robot, warm-started cycle after cycle (20 ms cadence, SRB step between cycles), "
                                           "OSQP restatement at eps 1e-5; the reference quotes 20-33 ms per cycle (README.md:50)"}
    sub = rec.slice(0, min(rec.B, 1024))
    r0 = cpu_port.solve_batch(sub, opts, nthreads=cores)
    nxt = records.next_cycle(sub, r0["w"][:, 12 * sub.N:12 * sub.N + 12])
    t = time.perf_counter()
    r1 = cpu_port.solve_batch(nxt, opts, warm=True, state=(r0["w"], r0["y"], r0["rho"]), nthreads=cores)
    tw = time.perf_counter() - t
    out["cpuP_warm"] = {"qps": sub.B / tw, "threads": cores, "robots": sub.B, "iters_mean": float(r1["iters"].mean()),
                        "solved_frac": float((r1["status"] == 1).mean()),
                        "what": "1 024 robots of the workload, second cycle warm-started from the first (20 ms later)"}
    return out


def _live_worker(job):
    """One process of the live reference arm: solve robots [lo, hi) of the seeded workload, cold start."""
    lo, hi, stress, reps = job
    from convex_mpc_b200 import records
    from oracle import live_reference
    live = live_reference.LiveReference(eps=1e-5)
    rec = records.random_records(65536, seed=65536, stress=stress).slice(lo, hi)
    live.solve(rec, 0)                                   # solver construction outside the timers, as in the reference
    t = time.perf_counter()
    for _ in range(reps):
        for b in range(rec.B):
            live.solve(rec, b)
    return time.perf_counter() - t


def run_reference_live(args, cores):
    """The reference's own CentroidalMPC (CasADi -> OSQP, eps 1e-5) on the first robots of the workload: once
    single-process and once one process per core (CPU-1 / CPU-P)."""
    import multiprocessing as mp
    t1 = _live_worker((0, 8, args.stress, 1))
    rate1 = 8 / t1
    step_s = min(2.0, 90.0 / max(1, args.steps + args.warmup))
    per_proc = int(max(1, rate1 * step_s))
    jobs = [(i * per_proc, (i + 1) * per_proc, args.stress, args.steps) for i in range(cores)]
    with mp.get_context("spawn").Pool(cores) as pool:
        ts = pool.map(_live_worker, jobs)
    el = max(ts)
    per_step = per_proc * cores
    val = per_step * args.steps / el
    sample = (f"{per_step} QPs per step ({per_proc} per process x {cores} processes, first robots of the 65536-robot "
              f"workload), cold start, the reference's CentroidalMPC.solve_QP (CasADi conic -> OSQP) at eps 1e-5; "
              f"single process: {rate1:.1f} QPs/s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "horizon": 16, "start": "cold", "sample_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                         "value_1process": rate1},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_reference(args, rank, world):
    """--impl reference: the CPU arm alone, rank 0 only.  The reference's own CasADi -> OSQP path when the casadi
    wheel and the reference tree are present (oracle/live_reference.py), else the restated port."""
    if rank != 0:
        return
    from convex_mpc_b200 import records
    from oracle import cpu_port, live_reference
    if live_reference.available():
        cores_l = os.cpu_count() or 1
        try:
            cores_l = len(os.sched_getaffinity(0))
        except Exception:
            pass
        return run_reference_live(args, cores_l)
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    eps = 1e-5
    opts = cpu_port.default_opts(eps_abs=eps, eps_rel=eps)
    # bounded sample per step: the whole run (warm-up + K steps) is sized to ~90 s of all-core work, at most 2 s per step
    probe = records.random_records(16, seed=65536, stress=args.stress)
    t = time.perf_counter(); cpu_port.solve_batch(probe, opts, nthreads=1); t1 = time.perf_counter() - t
    step_s = min(2.0, 90.0 / max(1, args.steps + args.warmup))
    per_step = int(max(cores, min(8192, (16 / t1) * cores * step_s)))
    rec = records.random_records(65536, seed=65536, stress=args.stress).slice(0, per_step)
    for _ in range(args.warmup):
        cpu_port.solve_batch(rec, opts, nthreads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        r = cpu_port.solve_batch(rec, opts, nthreads=cores)
    el = time.perf_counter() - t
    val = per_step * args.steps / el
    sample = (f"{per_step} QPs per step (first {per_step} of the 65536-robot workload), cold start, {cores} threads, "
              f"OSQP restatement eps={eps:g} on the reference's sparse QP; restated CPU baseline, not the reference binary")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "horizon": 16, "start": "cold", "sample_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "iters_mean": float(r["iters"].mean())},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    from convex_mpc_b200 import sharding
    rank, local_rank, world = sharding.env_rank_world()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from convex_mpc_b200 import _lib, records, roofline
    from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N = 16
    B = args.batch
    # weak scaling: every rank owns its own 64K robots of the (world * B)-robot job
    rec = records.random_records(B, seed=65536 + rank, stress=args.stress)
    max_stance = 4 * (int(np.floor(rec.duty * N)) + 1)        # periodic-gait bound on stance foot-steps
    traj = BatchedComTraj.from_records(rec, device=dev)
    mpc = CentroidalMPC(None, traj, verbose=False, mode=args.mode, max_stance=max_stance, device=dev, prepass=args.prepass)
    _lib.check(lib.cmpc_set_profile(mpc._h, 1))          # per-kernel CUDA events inside cmpc_solve

    if args.sweep:
        sweep(args, mpc, records, BatchedComTraj, CentroidalMPC, dev, max_stance)
        return

    def step():
        mpc.reset()                      # cold start: the whole build + solve is redone every step
        return mpc.solve_QP(None, traj)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    launches0 = lib.cmpc_launch_count()
    barrier()
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    kern_ms, k_pre, k_cert, k_con = [], [], [], []
    a_ms, b_ms, c_ms = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    for i in range(args.steps):
        step()
        kern_ms.append(mpc.kernel_ms)
        if lib.cmpc_last_kernel_ms3(mpc._h, ctypes.byref(a_ms), ctypes.byref(c_ms), ctypes.byref(b_ms)) == 0:
            k_pre.append(a_ms.value); k_cert.append(c_ms.value); k_con.append(b_ms.value)
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop()
    launches = lib.cmpc_launch_count() - launches0
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    value = world * B * args.steps / (total_ms * 1e-3)

    status = mpc._status.cpu().numpy()
    u_final = mpc._u.cpu().numpy()
    iters = mpc._iters.cpu().numpy()
    stats = mpc._stats.cpu().numpy()
    flops = roofline.batch_flops(stats, iters, N, prepass=args.prepass)
    flops_route = roofline.batch_flops(stats, iters, N, route_actual=True, prepass=args.prepass)
    f_pre, f_con = roofline.split_flops(stats, iters, N, prepass=args.prepass)
    kms = float(np.mean(kern_ms))
    pre_ms = float(np.mean(k_pre)) if k_pre else None
    con_ms = float(np.mean(k_con)) if k_con else None
    cert_ms = float(np.mean(k_cert)) if k_cert else None

    # ---- end-to-end through the host-buffer C-ABI call ------------------------------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hb = [pin(rec.x0), pin(rec.x_ref), pin(rec.r_foot), pin(rec.I_world), pin(rec.mass), pin(rec.t0)]
    out = (torch.empty(B, 12 * N, dtype=torch.float64).pin_memory(),
           torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
    mpc_h = CentroidalMPC(None, traj, verbose=False, mode=args.mode, max_stance=max_stance, max_batch=B, device=dev, prepass=args.prepass)

    def step_host():
        mpc_h._warm_host = 0
        return mpc_h.solve_host(*hb, rec.dt, rec.gait_hz, rec.duty, out=out)

    for _ in range(max(args.warmup, 3)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_val = world * B * args.steps / (e2e_ms * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in hb)
    d2h = sum(t.numel() * t.element_size() for t in out)
    assert (out[1].numpy() == status).all()

    # ---- second end-to-end figure: the whole cycle from state + command (cmpc_cycle_host), first-step forces out --------
    e2e_cycle = None
    try:
        g = records.random_cycle_inputs(B, seed=4242 + rank)
        cb = {k: pin(v) for k, v in g.items() if k != "hip"}
        out_c = (torch.empty(B, 12, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory(),
                 torch.empty(B, dtype=torch.int32).pin_memory())
        pos0 = g["pos_des"].copy()

        def step_cycle():
            mpc_h._warm_host = 0
            cb["pos_des"].numpy()[:] = pos0
            return mpc_h.cycle_host(cb["x0"], cb["R_wb"], cb["lever"], cb["cmd"], cb["t0"], cb["pos_des"], cb["I_world"], cb["mass"],
                                    rec.dt, g["hip"], gait_hz=rec.gait_hz, duty=rec.duty, first_step_only=True, out=out_c)

        for _ in range(max(args.warmup, 3)):
            step_cycle()
        barrier()
        t0c = time.perf_counter()
        for _ in range(args.steps):
            step_cycle()
        barrier()
        c_ms = max_over_ranks((time.perf_counter() - t0c) * 1e3)
        e2e_cycle = {"value": world * B * args.steps / (c_ms * 1e-3), "unit": UNIT, "ms_per_step": c_ms / args.steps,
                     "h2d_bytes_per_step": int(sum(cb[k].numel() * 8 for k in cb)), "d2h_bytes_per_step": int(B * (12 * 8 + 3 * 8 + 8)),
                     "solved_frac": float((out_c[1].numpy() == 1).mean()),
                     "api": "CentroidalMPC.cycle_host -> cmpc_cycle_host: state + command in (408 B/robot), ComTraj.generate_traj on the "
                            "device, first-step forces U_opt[:,0] out (96 B/robot); cold start every step"}
    except Exception as e:
        e2e_cycle = {"error": repr(e)}

    # ---- statistics gathered across ranks (the only collective of the run) ----------------------
    recs = sharding.gather_stats(sharding.local_stats(status, iters, stats, total_ms, flops, flops_route), device=dev)
    summ = sharding.reduce_stats(recs)

    line = None
    f64 = ctypes.c_double(); smem = ctypes.c_double()
    if rank == 0:
        _lib.check(lib.cmpc_microbench(local_rank, ctypes.byref(f64), ctypes.byref(smem)))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 (of fallback)"
        ach_tf = flops / (kms * 1e-3) / 1e12
        alg_bytes = roofline.bytes_per_qp(N) * B
        # DRAM traffic per launch from the committed ncu capture of the same kernel (per-robot figure x robots)
        traffic, traffic_src = None, None
        try:
            import glob, re
            caps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_*_dram_traffic.json")),
                          key=lambda f: [int(x) for x in re.findall(r"\d+", os.path.basename(f))])
            tr = json.load(open(caps[-1]))                      # the newest committed capture
            if N == 16 and args.mode == "active_set":
                traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["robots"] * B
                traffic_src = f"profiles/{os.path.basename(caps[-1])} (ncu dram__bytes_read.sum + dram__bytes_write.sum, scaled per robot)"
        except Exception:
            pass
        pre_name = {4: "wrench_pdas_kernel", 3: "riccati2_lockstep_kernel", 2: "riccati2_kernel", 1: "riccati_kernel"}.get(args.prepass)
        kernels = []
        if pre_name and pre_ms:
            kernels.append({"name": pre_name, "ms": pre_ms, "algorithmic_flops": f_pre,
                            "achieved": f_pre / (pre_ms * 1e-3) / 1e12, "frac": f_pre / (pre_ms * 1e-3) / 1e12 / f64.value,
                            "robots_finished": int(np.isin(stats[:, 7].astype(int), (4, 5)).sum())})
        if args.prepass == 4 and cert_ms:
            # the certificate pass: 0.7 kFLOP per stage and robot (roofline.py), HBM-bound rather than FP64-bound: X, u, y, x_ref,
            # r_foot in (76 N doubles), co-states and box rows of y out (24 N doubles) per robot
            nfin = int(np.isin(stats[:, 7].astype(int), (4, 5)).sum())
            f_cert = nfin * N * 0.7e3
            b_cert = nfin * 8.0 * (76 * N + 24 * N)
            kernels[-1]["algorithmic_flops"] = f_pre - f_cert
            kernels[-1]["achieved"] = (f_pre - f_cert) / (pre_ms * 1e-3) / 1e12
            kernels[-1]["frac"] = kernels[-1]["achieved"] / f64.value
            kernels.append({"name": "wrench_certificate_kernel", "ms": cert_ms, "algorithmic_flops": f_cert,
                            "achieved": f_cert / (cert_ms * 1e-3) / 1e12, "frac": f_cert / (cert_ms * 1e-3) / 1e12 / f64.value,
                            "bound": "hbm", "algorithmic_bytes": b_cert, "hbm_gbs": b_cert / (cert_ms * 1e-3) / 1e9,
                            "hbm_frac": b_cert / (cert_ms * 1e-3) / 1e9 / hbm_peak, "robots_finished": nfin})
        if con_ms is not None:
            # route 4: the condensed kernel runs on the hand-overs WHILE the certificate kernel runs (two streams); what is
            # timed here is what is left of it after the certificates are done
            kernels.append({"name": "solve_fast_kernel" + (" (remainder after the overlap with the certificate kernel)" if args.prepass == 4 else ""), "ms": con_ms, "algorithmic_flops": f_con,
                            "achieved": f_con / (max(con_ms, 1e-6) * 1e-3) / 1e12,
                            "frac": f_con / (max(con_ms, 1e-6) * 1e-3) / 1e12 / f64.value,
                            "robots_finished": int((~np.isin(stats[:, 7].astype(int), (4, 5))).sum())})
        dom = max(kernels, key=lambda k_: k_["ms"]) if kernels else None
        ach_route = flops_route / (kms * 1e-3) / 1e12
        roof = {"bound": "fp64_fma",
                "kernel": (dom["name"] if dom else "solve_fast_kernel") + " (dominant kernel of one cmpc_solve call)",
                "achieved": dom["achieved"] if dom else ach_route, "peak": f64.value, "unit": "TFLOP/s",
                "frac": dom["frac"] if dom else ach_route / f64.value,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel_ms": dom["ms"] if dom else kms, "algorithmic_flops_per_launch": dom["algorithmic_flops"] if dom else flops_route,
                "kernels": kernels,
                "whole_solve": {"ms": kms, "algorithmic_flops": flops_route, "achieved": ach_route, "frac": ach_route / f64.value},
                "frac_condensed_route_figure": ach_tf / f64.value,
                "note": "frac = algorithmic flops of the route every robot really took (convex_mpc_b200/roofline.py: 6.8 kFLOP per stage and "
                        "sweep + 0.7 kFLOP per stage for the certificate on the Riccati route; Cholesky / active-set / ADMM counts on the "
                        "condensed route) / live CUDA-event duration of that kernel / measured DFMA peak; "
                        "frac_condensed_route_figure charges every robot the condensed route's flops (SURVEY 8d per-unit figure, round-1 comparable)",
                "peak_source": "cmpc_microbench DFMA stream measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "smem_gbs_measured": smem.value,
                "hbm": {"achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                        "algorithmic_bytes_per_launch": alg_bytes}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "p50_batch_ms": float(np.median(per_step)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "horizon": N, "mode": args.mode, "start": "cold",
                       "stress": args.stress, "max_stance": max_stance,
                       "l2": f"inputs {B * 3280 / 1e6:.0f} MB + outputs {B * 6216 / 1e6:.0f} MB per step exceed the 126 MB L2",
                       "parallelism": f"batch sharded over {world} GPU(s), no per-step communication"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "api": "CentroidalMPC.solve_host -> cmpc_solve_host (pinned host buffers)"},
            "e2e_cycle": e2e_cycle,
            "gpu_launches": int(launches),
            "roofline": roof,
            "solver": {k: summ[k] for k in ("solved", "max_iter", "inaccurate", "failed", "path_unconstrained",
                                            "path_active_set", "path_admm", "path_admm_polish", "path_riccati", "path_wrench_as", "r_prim_max", "r_dual_max")},
        }
        line["solver"]["as_iters_mean"] = summ["as_iters_sum"] / max(summ["qps_count"], 1)
        line["solver"]["n_free_mean"] = summ["n_free_sum"] / max(summ["qps_count"], 1)

    # ---- extras on rank 0 at N=1: ADMM-only mode, warm start, CPU baseline ------------------------
    if rank == 0 and world == 1 and not args.no_extras:
        extra = {}
        try:
            sub = rec.slice(0, min(B, 16384))
            tr = BatchedComTraj.from_records(sub, device=dev)
            m2 = CentroidalMPC(None, tr, verbose=False, mode="admm", eps_abs=1e-5, eps_rel=1e-5, max_iter=4000,
                               max_stance=max_stance, device=dev)
            m2.solve_QP(None, tr)
            ts = []
            for _ in range(3):
                m2.reset(); m2.solve_QP(None, tr); ts.append(m2.kernel_ms)
            it2 = m2._iters.cpu().numpy()
            st2 = m2._stats.cpu().numpy()
            nblk = np.ceil(st2[:, 3] / 8.0)
            # algorithmic shared-memory traffic of the ADMM iterations: two passes over the block-packed
            # inverse factor per iteration (x~ = W^T (W rhs)), 512 B per 8x8 block
            smem_bytes = float((it2 * 2.0 * nblk * (nblk + 1) / 2.0 * 512.0).sum())
            t_admm = float(np.median(ts)) * 1e-3
            extra["admm_mode"] = {"value": sub.B / t_admm, "unit": UNIT, "batch": sub.B,
                                  "eps": 1e-5, "iters_mean": float(it2.mean()),
                                  "solved_frac": float((m2._status.cpu().numpy() == 1).mean()),
                                  "roofline": {"bound": "shared_memory", "achieved": smem_bytes / t_admm / 1e9,
                                               "peak": smem.value, "unit": "GB/s", "frac": smem_bytes / t_admm / 1e9 / smem.value,
                                               "note": "iterations only in the numerator, whole kernel (incl. 1-3 factorisations "
                                                       "and build) in the denominator; peak = smem read stream measured in this run"}}
            # warm start: second solve of the same batch from the previous solution
            mpc.solve_QP(None, traj)
            ts = []
            for _ in range(3):
                mpc.solve_QP(None, traj); ts.append(mpc.kernel_ms)
            extra["warm_start"] = {"value": B / (np.median(ts) * 1e-3), "unit": UNIT}
        except Exception as e:      # extras never break the contract line
            extra["error"] = repr(e)
        line["extra"] = extra
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(rec)
        try:
            line["solver"]["l2_check"] = l2_check(rec, u_final, 16)
        except Exception as e:
            line["solver"]["l2_check"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sweep(args, mpc, records, BatchedComTraj, CentroidalMPC, dev, max_stance):
    import torch
    out = []
    for B in [1, 2, 4, 8, 16, 32, 64, 148, 296, 1024, 4096, 16384, 65536, 262144]:
        rec = records.random_records(B, seed=65536, stress=args.stress)
        tr = BatchedComTraj.from_records(rec, device=dev)
        m = CentroidalMPC(None, tr, verbose=False, mode=args.mode, max_stance=max_stance, device=dev, prepass=args.prepass)
        for _ in range(3):
            m.reset(); m.solve_QP(None, tr)
        ts = []
        for _ in range(7 if B <= 65536 else 3):
            m.reset(); m.solve_QP(None, tr); ts.append(m.kernel_ms)
        t = float(np.median(ts))
        out.append({"batch": B, "p50_ms": t, "qps": B / t * 1e3,
                    "solved_frac": float((m._status == 1).float().mean().item())})
        print(json.dumps(out[-1]), flush=True)
        del m, tr
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_{args.mode}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
