"""Device-resident timing of cmpc_solve with the wrench-space PDAS pre-pass against the round-1 route (development aid)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import records  # noqa: E402
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    cases = [(16, 0.0), (16, 0.05), (16, 0.3), (32, 0.0), (48, 0.0)]
    if len(sys.argv) > 2:
        cases = [(int(sys.argv[2]), float(sys.argv[3]))]
    for N, stress in cases:
        b = B if N == 16 else B // 4
        rec = records.random_records(b, N=N, seed=65536, stress=stress)
        traj = BatchedComTraj.from_records(rec, device="cuda:0")
        for pp in (4, 3):
            ms_ = 4 * (int(np.floor(rec.duty * N)) + 1)
            mpc = CentroidalMPC(None, traj, verbose=False, max_stance=ms_, prepass=pp)
            for _ in range(2):
                mpc.reset(); mpc.solve_QP(None, traj)
            ts = []
            for _ in range(5):
                mpc.reset(); mpc.solve_QP(None, traj); ts.append(mpc.kernel_ms)
            st = mpc.last_stats.cpu().numpy()
            t = float(np.median(ts))
            print(json.dumps({"N": N, "B": b, "stress": stress, "prepass": pp, "ms": round(t, 3), "Mqps": round(b / t / 1e3, 3),
                              "paths": np.bincount(st[:, 7].astype(int), minlength=6).tolist(),
                              "sweeps_mean": float(1 + st[:, 6][np.isin(st[:, 7], (4, 5))].mean()) if pp == 4 else None,
                              "solved": float((mpc._status == 1).float().mean().item())}), flush=True)
            del mpc


if __name__ == "__main__":
    main()
