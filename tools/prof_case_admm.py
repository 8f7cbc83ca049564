"""One small ADMM-mode solve for ncu: python tools/prof_case_admm.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import records
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
rec = records.random_records(B, seed=65536, stress=0.0)
traj = BatchedComTraj.from_records(rec, device="cuda:0")
mpc = CentroidalMPC(None, traj, verbose=False, max_stance=40, mode="admm", eps_abs=1e-5, eps_rel=1e-5, max_iter=4000)
for _ in range(3):
    mpc.reset(); mpc.solve_QP(None, traj)
print("kernel ms", mpc.kernel_ms, "ok", int((mpc._status == 1).sum().item()), "/", B, "iters", float(mpc._iters.float().mean().item()))
