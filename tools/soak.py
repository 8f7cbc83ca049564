"""Robustness soak: several seeds of the headline workload (and mildly disturbed variants) through the default path;
every robot must come back solved with a clean certificate.  python tools/soak.py [B] [seeds]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import records
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
S = int(sys.argv[2]) if len(sys.argv) > 2 else 6
out = []
for stress in (0.0, 0.05):
    for seed in range(1, S + 1):
        rec = records.random_records(B, seed=1000 * seed + 7, stress=stress)
        traj = BatchedComTraj.from_records(rec, device="cuda:0")
        mpc = CentroidalMPC(None, traj, verbose=False, max_stance=40)
        sol = mpc.solve_QP(None, traj)
        st = sol["status"].cpu().numpy(); stats = sol["stats"].cpu().numpy(); u1 = sol["u"].clone()
        sol2 = mpc.solve_QP(None, traj)                         # warm-started second cycle on the same data
        st2 = sol2["status"].cpu().numpy()
        exact = ~np.isin(stats[:, 7], (2, 3))
        du = float((sol2["u"] - u1).abs().amax(dim=(1, 2)).cpu().numpy()[exact].max())
        row = {"stress": stress, "seed": seed, "robots": B, "solved": int((st == 1).sum()), "other_status": np.unique(st[st != 1]).tolist(),
               "paths": np.bincount(stats[:, 7].astype(int), minlength=5).tolist(), "r_prim_max": float(stats[:, 0].max()),
               "r_dual_max": float(stats[:, 1].max()), "warm_solved": int((st2 == 1).sum()), "warm_vs_cold_max_du": du, "ms": mpc.kernel_ms}
        out.append(row); print(json.dumps(row), flush=True)
        del mpc, traj
json.dump(out, open(os.path.join("gpurun_out", "soak.json"), "w"), indent=1)
