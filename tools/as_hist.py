"""Histogram of active-set sizes / iterations on the synthetic workload: python tools/as_hist.py [B] [stress]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import records
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for stress in ([float(sys.argv[2])] if len(sys.argv) > 2 else [0.0, 0.3, 1.0]):
    rec = records.random_records(B, seed=65536, stress=stress)
    traj = BatchedComTraj.from_records(rec, device="cuda:0")
    mpc = CentroidalMPC(None, traj, verbose=False, max_stance=40)
    mpc.solve_QP(None, traj)
    st = mpc.last_stats.cpu().numpy()
    path = st[:, 7].astype(int); k = st[:, 4].astype(int); it = st[:, 6].astype(int)
    print(f"stress {stress}: paths {np.bincount(path, minlength=4).tolist()}")
    sel = path >= 1
    print("  n_active percentiles (active-set QPs) 50/90/99/max:", np.percentile(k[sel], [50, 90, 99, 100]).tolist())
    print("  n_active hist (0..40+):", np.bincount(np.minimum(k[sel], 40), minlength=41).tolist())
    print("  as_iters percentiles 50/90/99/max:", np.percentile(it[sel], [50, 90, 99, 100]).tolist())
