"""Record / replay of closed-loop states (BASELINE configs[1]: 1 024 Go2 robots, mixed forward / lateral 0.4 m/s / yaw
4 rad/s commands, one B200).

    python tools/replay.py record  [B=1024] [cycles=500] [file=gpurun_out/replay_1024x500.npz]
    python tools/replay.py replay  [file] [out=gpurun_out/replay.json]

record: the robots are driven cycle by cycle -- solve_QP on the GPU (warm-started), the single-rigid-body step between
cycles on the host (records.next_cycle, the stand-in for MuJoCo) -- and every cycle's record and forces are written in
the format of records.save_cycles.  replay: the recorded cycles are fed to the solver in order, batch B, warm-started
from cycle c-1, and the forces are compared with the recorded ones; the replay time per cycle is the configs[1] figure.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import records  # noqa: E402
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC  # noqa: E402


def record(B, C, path):
    rng = np.random.default_rng(1024)
    rec = records.random_records(B, seed=1024, stress=0.0)
    rec = records.retarget(rec, rng.uniform(-0.8, 0.8, B), rng.uniform(-0.4, 0.4, B), rng.uniform(-4.0, 4.0, B))
    mpc, cyc, us = None, [], []
    seg = rng.integers(25, 100, B)            # cycles until the next command change (0.5 - 2 s)
    for c in range(C):
        change = (c % seg) == 0
        if c and change.any():
            new = records.retarget(rec, rng.uniform(-0.8, 0.8, B), rng.uniform(-0.4, 0.4, B), rng.uniform(-4.0, 4.0, B))
            rec.x_ref[change] = new.x_ref[change]
        traj = BatchedComTraj.from_records(rec, device="cuda:0")
        if mpc is None:
            mpc = CentroidalMPC(None, traj, verbose=False, max_stance=4 * (int(np.floor(rec.duty * rec.N)) + 1))
        sol = mpc.solve_QP(None, traj)
        assert (sol["status"].cpu().numpy() == 1).all(), c
        u = mpc._u.cpu().numpy()
        cyc.append(rec); us.append(u.copy())
        rec = records.next_cycle(rec, u[:, :12])
    records.save_cycles(path, cyc, us)
    return cyc, us


def replay(path, out):
    cyc, us = records.load_cycles(path)
    B, C = cyc[0].B, len(cyc)
    trajs = [BatchedComTraj.from_records(r, device="cuda:0") for r in cyc]
    mpc = CentroidalMPC(None, trajs[0], verbose=False, max_stance=4 * (int(np.floor(cyc[0].duty * cyc[0].N)) + 1))
    worst, kms = 0.0, []
    torch.cuda.synchronize()
    t = time.perf_counter()
    for c in range(C):
        sol = mpc.solve_QP(None, trajs[c])
        kms.append(mpc.kernel_ms)
        if us is not None and (c % 25 == 0 or c == C - 1):
            worst = max(worst, float(np.abs(mpc._u.cpu().numpy() - us[c]).max()))
    torch.cuda.synchronize()
    el = time.perf_counter() - t
    res = {"robots": B, "cycles": C, "ms_per_cycle": el / C * 1e3, "kernel_ms_p50": float(np.median(kms)), "qps_per_s": B * C / el,
           "max_abs_force_difference_to_recording_N": worst, "file": os.path.basename(path)}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))
    return res


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "record"
    if mode == "record":
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
        C = int(sys.argv[3]) if len(sys.argv) > 3 else 500
        path = sys.argv[4] if len(sys.argv) > 4 else os.path.join("gpurun_out", f"replay_{B}x{C}.npz")
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        record(B, C, path)
        print("recorded", path, os.path.getsize(path) // 1024, "KiB")
    else:
        path = sys.argv[2]
        replay(path, sys.argv[3] if len(sys.argv) > 3 else os.path.join("gpurun_out", "replay.json"))
