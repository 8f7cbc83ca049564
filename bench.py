#!/usr/bin/env python
"""bench.py -- the contract benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--mode M]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): MPC QPs solved/sec at horizon N=16, batch 64K per GPU; p50 batch latency.
Workload: BASELINE configs[2] -- 65 536 randomized Go2 trot states + references per GPU, cold start
(every step re-solves the batch from scratch; inputs of one batch, 213 MB, exceed the 126 MB L2).
One "step" = one batched ``solve_QP`` (contact table + dynamics + condensed QP build + solve, fused).

  value      QPs/s with inputs resident in HBM (CUDA events, barrier + sync on both sides, max over ranks)
  e2e        the same through the host-buffer C-ABI call ``cmpc_solve_host`` (pinned host memory in,
             forces/status out; H2D and D2H inside the timed region)
  roofline   fused solve kernel vs the FP64-FMA roofline (peak measured on this GPU by the library's
             DFMA micro-benchmark -- MEASURED_PEAKS.json has no FP64 entry) + achieved HBM GB/s vs
             MEASURED_PEAKS.json
  cpu_baseline   the oracle's C port of the reference CPU path (OSQP restatement) on the host cores

``--impl reference`` times that CPU port alone (the reference itself needs CasADi/OSQP/Pinocchio,
none of which is installable here -- DESIGN.md section 7).
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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="robots per GPU")
    ap.add_argument("--mode", default="active_set", choices=["active_set", "admm"])
    ap.add_argument("--stress", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--prepass", type=int, default=3, help="Riccati pre-pass ahead of the condensed kernel: 0 off, 1 v1, 2 v2, 3 v2 lock-step (default)")
    ap.add_argument("--sweep", action="store_true", help="batch-size sweep 1..262144 -> gpurun_out/sweep.json")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
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
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
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
    return {"value": n / tp, "unit": UNIT, "cores": cores, "kind": "port",
            "value_1thread": rate1,
            "iters_mean": float(r["iters"].mean()), "solved_frac": float((r["status"] == 1).mean()),
            "sample": f"first {n} QPs of the same workload, cold start, OSQP restatement (oracle/osqp_port.c) on the "
                      f"reference's sparse 384-var QP, reference OPTS but eps_abs=eps_rel={eps:g}, {cores} threads "
                      f"({tp:.1f} s); single-thread probe {probe.B} QPs ({t1:.1f} s). Restated CPU baseline, not the "
                      f"reference binary (CasADi/OSQP not installable)"}


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
    kern_ms = []
    for i in range(args.steps):
        step()
        kern_ms.append(mpc.kernel_ms)
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop()
    launches = lib.cmpc_launch_count() - launches0
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[-1]))
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    value = world * B * args.steps / (total_ms * 1e-3)

    status = mpc._status.cpu().numpy()
    iters = mpc._iters.cpu().numpy()
    stats = mpc._stats.cpu().numpy()
    flops = roofline.batch_flops(stats, iters, N)
    flops_route = roofline.batch_flops(stats, iters, N, route_actual=True)
    kms = float(np.mean(kern_ms))

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
            caps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_v*_dram_traffic.json")),
                          key=lambda f: [int(x) for x in re.findall(r"\d+", os.path.basename(f))])
            tr = json.load(open(caps[-1]))                      # the newest committed capture
            if N == 16 and args.mode == "active_set":
                traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["robots"] * B
                traffic_src = f"profiles/{os.path.basename(caps[-1])} (ncu dram__bytes_read.sum + dram__bytes_write.sum, scaled per robot)"
        except Exception:
            pass
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
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64_fma", "kernel": "riccati2_lockstep_kernel + solve_fast_kernel (one cmpc_solve call)" if args.prepass else "solve_fast_kernel", "achieved": ach_tf, "peak": f64.value,
                         "unit": "TFLOP/s", "frac": ach_tf / f64.value, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel_ms": kms, "algorithmic_flops_per_launch": flops,
                         "frac_route_actual": flops_route / (kms * 1e-3) / 1e12 / f64.value,
                         "note": "frac counts every robot at the condensed route's flops (SURVEY 8d per-unit figure); "
                                 "frac_route_actual counts robots finished by the Riccati pre-pass at that sweep's (4.5x smaller) flops; "
                                 "kernel_ms = pre-pass + condensed kernel of one cmpc_solve call",
                         "peak_source": "cmpc_microbench DFMA stream measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "smem_gbs_measured": smem.value,
                         "hbm": {"achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                                 "algorithmic_bytes_per_launch": alg_bytes}},
            "solver": {k: summ[k] for k in ("solved", "max_iter", "inaccurate", "failed", "path_unconstrained",
                                            "path_active_set", "path_admm", "path_admm_polish", "path_riccati", "r_prim_max", "r_dual_max")},
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
