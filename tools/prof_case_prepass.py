"""One small solve with the Riccati pre-pass for ncu: python tools/prof_case_prepass.py [B] [stress] [max_stance] [version]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import records  # noqa: E402
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
stress = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
ms = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ver = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rec = records.random_records(B, seed=65536, stress=stress)
traj = BatchedComTraj.from_records(rec, device="cuda:0")
mpc = CentroidalMPC(None, traj, verbose=False, max_stance=ms, prepass=ver)
for _ in range(3):
    mpc.reset()
    mpc.solve_QP(None, traj)
print("solve_time ms", mpc.solve_time, "status ok", int((mpc._status == 1).sum().item()), "/", B)
