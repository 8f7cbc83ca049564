"""Device-resident closed loop (BASELINE configs[1] shape): B robots x C MPC cycles,
    generate_traj -> contact table + solve_QP (warm-started) -> srb_step
all on the GPU (cmpc_generate_traj, cmpc_contact_table, cmpc_solve, cmpc_srb_step); the host only enqueues.

    python tools/closed_loop.py [B=1024] [cycles=500] [out=gpurun_out/closed_loop.json] [shift]

``shift``: warm start from the previous working set shifted by one horizon stage instead of the
unshifted previous solution.  With CMPC_PREPASS_MIN_BATCH=1 the Riccati-sweep kernel also serves batches below 2 048
robots; ``sweeps_mean_sampled`` then tells how many sweeps a warm-started cycle needs.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import com_trajectory as ct, records  # noqa: E402
from convex_mpc_b200.centroidal_mpc import CentroidalMPC  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
C = int(sys.argv[2]) if len(sys.argv) > 2 else 500
OUT = sys.argv[3] if len(sys.argv) > 3 else os.path.join("gpurun_out", "closed_loop.json")
SHIFT = len(sys.argv) > 4 and sys.argv[4] == "shift"
dev = torch.device("cuda:0")
rng = np.random.default_rng(1024)
HZ, DUTY, N, MPC_DT = 3.0, 0.6, 16, 0.02
gait = ct.Gait(HZ, DUTY)
dt = gait.gait_period / N
hip = np.array([[0.1934, 0.0465, 0], [0.1934, -0.0465, 0], [-0.1934, 0.0465, 0], [-0.1934, -0.0465, 0]])
stance = np.array([[records.HIP_X, records.HIP_Y, 0], [records.HIP_X, -records.HIP_Y, 0],
                   [-records.HIP_X, records.HIP_Y, 0], [-records.HIP_X, -records.HIP_Y, 0]])

# start: standing at height 0.27 with random yaw, per-robot piecewise-constant commands (fwd / lateral / yaw rate)
yaw = rng.uniform(-np.pi, np.pi, B)
x = np.zeros((B, 12)); x[:, 0:2] = rng.uniform(-5, 5, (B, 2)); x[:, 2] = 0.27; x[:, 5] = yaw
R = records._rot_zyx(x[:, 3], x[:, 4], x[:, 5])
lever = np.zeros((B, 4, 3))
for leg in range(4):
    lever[:, leg, 0] = np.cos(yaw) * stance[leg, 0] - np.sin(yaw) * stance[leg, 1]
    lever[:, leg, 1] = np.sin(yaw) * stance[leg, 0] + np.cos(yaw) * stance[leg, 1]
    lever[:, leg, 2] = -0.27
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
state = ct.RobotState(t(x), t(np.swapaxes(R, 1, 2).copy()), t(lever), t(np.full(B, records.GO2_MASS)),
                      t(np.einsum("bij,j,bkj->bik", R, records.GO2_I_BODY, R)))
cmd = [t(rng.uniform(-0.8, 0.8, B)), t(rng.uniform(-0.4, 0.4, B)), t(np.full(B, 0.27)), t(rng.uniform(-4.0, 4.0, B))]   # BASELINE configs[1]: forward, lateral 0.4 m/s, yaw 4 rad/s

traj = ct.ComTraj(state, hip_offset=hip, device=dev)
traj.generate_traj(state, gait, 0.0, *cmd, dt)
mpc = CentroidalMPC(None, traj, verbose=False, max_stance=40, warm_shift=SHIFT)
sweeps = []
nxt = None
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
solved = 0
worst_rp = worst_rd = 0.0
paths = np.zeros(5)
WARM = 10
nchecks = 0
for c in range(C + WARM):
    if c == WARM:
        torch.cuda.synchronize(); ev[0].record()
    traj.generate_traj(state, gait, c * MPC_DT, *cmd, dt)
    sol = mpc.solve_QP(None, traj)
    nxt2 = ct.srb_step(state, traj, mpc._u, MPC_DT, records.GO2_I_BODY, stance, out=nxt)
    state, nxt = nxt2, state
    if c >= WARM and (c % 50 == 0 or c == C + WARM - 1):
        st = sol["stats"].cpu().numpy(); nchecks += 1
        solved += int((sol["status"].cpu().numpy() == 1).sum()); worst_rp = max(worst_rp, st[:, 0].max()); worst_rd = max(worst_rd, st[:, 1].max())
        paths += np.bincount(st[:, 7].astype(int), minlength=6)[:5]
        rs = np.isin(st[:, 7].astype(int), (4, 5))
        if rs.any():
            sweeps.append(float(1 + st[rs, 6].mean()))
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1])
xf = state.x.cpu().numpy()
res = {"robots": B, "cycles": C, "ms_total": ms, "ms_per_cycle": ms / C, "cycles_per_s": C / ms * 1e3, "qps_per_s": B * C / ms * 1e3,
       "checked_cycles_all_solved": bool(solved == B * nchecks), "warm_shift": SHIFT,
       "sweeps_mean_sampled": float(np.mean(sweeps)) if sweeps else None,
       "r_prim_max": worst_rp, "r_dual_max": worst_rd, "paths_sampled": paths.tolist(),
       "height_mean": float(xf[:, 2].mean()), "height_min": float(xf[:, 2].min()), "speed_mean": float(np.linalg.norm(xf[:, 6:8], axis=1).mean()),
       "note": "eager: includes the Python enqueue overhead of three C-ABI calls + solution dict per cycle and solve_QP's two stream synchronisations"}

# ---- the same loop captured once in a CUDA graph (two cycles: the state buffers ping-pong) and replayed
t0 = torch.full((B,), (C + WARM) * MPC_DT, dtype=torch.float64, device=dev)
cmd4 = torch.stack(cmd, dim=1).contiguous()
sA, sB = state, nxt
side = torch.cuda.Stream(device=dev)
side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    def two_cycles():
        for a, b in ((sA, sB), (sB, sA)):
            traj.enqueue_generate(a, gait, t0, cmd4, dt)
            mpc.enqueue(traj)
            ct.srb_step(a, traj, mpc._u, MPC_DT, records.GO2_I_BODY, stance, out=b)
            t0.add_(MPC_DT)
    for _ in range(3):
        two_cycles()                     # warm-up on the capture stream: every allocation happens here
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        two_cycles()
    reps = C // 2
    for _ in range(5):
        g.replay()
    side.synchronize()
    ev[2].record(side)
    for _ in range(reps):
        g.replay()
    ev[3].record(side)
    side.synchronize()
msg = ev[2].elapsed_time(ev[3])
st = mpc._stats.cpu().numpy(); stt = mpc._status.cpu().numpy(); xg = sA.x.cpu().numpy()
res["cuda_graph"] = {"cycles": 2 * reps, "ms_total": msg, "ms_per_cycle": msg / (2 * reps), "cycles_per_s": 2 * reps / msg * 1e3,
                     "qps_per_s": B * 2 * reps / msg * 1e3, "last_cycle_all_solved": bool((stt == 1).all()),
                     "r_prim_max": float(st[:, 0].max()), "r_dual_max": float(st[:, 1].max()),
                     "height_mean": float(xg[:, 2].mean()), "sim_time_s": float(t0[0].item()),
                     "note": "generate_traj + contact table + fused solve + SRB step of two consecutive cycles captured in one CUDA "
                             "graph (device-side time stamp, ping-pong state buffers) and replayed; no host work per cycle"}
os.makedirs(os.path.dirname(OUT) or ".", exist_ok=True)
json.dump(res, open(OUT, "w"), indent=1)
print(json.dumps(res))
