"""BASELINE configs[3]: horizon sweep N = 16/32/48 (192/384/576 condensed variables) at batch 16384:
QPs/s, status counts and the force error against the oracle's exact optimum on a sample.
    python tests/horizon_sweep.py [B] > gpurun_out/horizon_sweep.json"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))   # lives under tests/: it checks a sample against the oracle
from convex_mpc_b200 import records
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC
from helpers import force_error, oracle_solution

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
out = []
for N in (16, 32, 48):
    rec = records.random_records(B, N=N, seed=16384)
    traj = BatchedComTraj.from_records(rec, device="cuda:0")
    ms = 4 * (int(np.floor(rec.duty * N)) + 1)
    mpc = CentroidalMPC(None, traj, verbose=False, max_stance=ms)
    for _ in range(2):
        mpc.reset(); mpc.solve_QP(None, traj)
    ts = []
    for _ in range(5):
        mpc.reset(); mpc.solve_QP(None, traj); ts.append(mpc.kernel_ms)
    u = mpc._u.cpu().numpy(); st = mpc._status.cpu().numpy(); stats = mpc._stats.cpu().numpy()
    errs = []
    for b in range(0, B, max(1, B // 12)):
        o = oracle_solution(rec, b)
        errs.append(force_error(u[b], o["sol"]["U"]))
    t = float(np.median(ts))
    row = {"horizon": N, "condensed_vars": 12 * N, "batch": B, "max_stance": ms, "kernel_ms_p50": t, "qps": B / t * 1e3,
           "solved_frac": float((st == 1).mean()), "paths": np.bincount(stats[:, 7].astype(int), minlength=5).tolist(),
           "n_free_mean": float(stats[:, 3].mean()), "r_prim_max": float(stats[:, 0].max()), "r_dual_max": float(stats[:, 1].max()),
           "force_err_abs_max_N": float(max(e[0] for e in errs)), "force_err_vs_tolerance_max": float(max(e[1] for e in errs)),
           "oracle_samples": len(errs),
           "factor_location": "shared memory" if N == 16 else "L2-resident global scratch (does not fit 227 KB of shared memory)"}
    out.append(row)
    print(json.dumps(row), flush=True)
    del mpc, traj
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "horizon_sweep.json"), "w"), indent=1)
