"""Quick device-resident timing sweep (development aid; bench.py is the contract benchmark)."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import _lib, records  # noqa: E402
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC  # noqa: E402


def main():
    lib = _lib.load()
    f = ctypes.c_double(); s = ctypes.c_double()
    _lib.check(lib.cmpc_microbench(0, ctypes.byref(f), ctypes.byref(s)))
    print(json.dumps({"fp64_tflops": f.value, "smem_gbs": s.value}))
    sizes = [int(a) for a in sys.argv[1:]] or [1, 148, 1024, 8192, 65536]
    for stress in (0.0, 0.3):
        for B in sizes:
            rec = records.random_records(B, seed=65536, stress=stress)
            traj = BatchedComTraj.from_records(rec, device="cuda:0")
            for ms in (64, 40):
                mpc = CentroidalMPC(None, traj, verbose=False, max_stance=ms)
                for _ in range(2):
                    mpc.reset()
                    mpc.solve_QP(None, traj)
                ts = []
                for _ in range(5):
                    mpc.reset()
                    mpc.solve_QP(None, traj)
                    ts.append(mpc.solve_time)
                st = mpc.last_stats.cpu().numpy()
                path = np.bincount(st[:, 7].astype(int), minlength=5).tolist()
                t = float(np.median(ts))
                print(json.dumps({"B": B, "stress": stress, "max_stance": ms, "ms": round(t, 3),
                                  "qps": round(B / t * 1e3, 1), "paths": path,
                                  "as_iters_mean": float(st[:, 6].mean())}))
                del mpc


if __name__ == "__main__":
    main()
