"""Synthetic Go2 MPC records (the inputs of one ``solve_QP`` call per robot), batched.

The reference produces these per cycle in ``com_trajectory.py:37-211`` from Pinocchio/MuJoCo state,
neither of which exists here, so the workload is generated: SURVEY.md section 8(d) configs #1-#5.  A record
is the *raw* input of the hot path (what ``ComTraj`` holds before ``_continuousDynamics``):

    x0      (B, 12)        [p, rpy, v, omega] world frame         com_trajectory.py:37
    x_ref   (B, 12, N)     rows [p; rpy; v; omega], col i <-> (i+1) dt   com_trajectory.py:15-25,84-103
    r_foot  (B, 4, 3, N)   lever arms CoM->foot, world, legs FL FR RL RR  com_trajectory.py:204-207
    I_world (B, 3, 3)      centroidal inertia in world frame      com_trajectory.py:40
    mass    (B,)                                                  com_trajectory.py:39
    t0      (B,)           time_now fed to the gait schedule      com_trajectory.py:106
    dt, gait_hz, duty      scalars                                test_MPC.py:50-52,67

Host-side NumPy only: generating synthetic inputs is not part of the hot path.
"""
from dataclasses import dataclass

import numpy as np

GO2_MASS = 15.02                      # placeholder Go2 values (URDF not in the reference tree)
GO2_I_BODY = np.array([0.11, 0.33, 0.38])
HIP_X, HIP_Y = 0.1934, 0.142
PHASE_OFFSET = np.array([0.5, 0.0, 0.0, 0.5])   # gait.py:8


@dataclass
class Records:
    x0: np.ndarray
    x_ref: np.ndarray
    r_foot: np.ndarray
    I_world: np.ndarray
    mass: np.ndarray
    t0: np.ndarray
    dt: float
    gait_hz: float
    duty: float
    N: int

    @property
    def B(self):
        return self.x0.shape[0]

    def slice(self, lo, hi):
        return Records(self.x0[lo:hi], self.x_ref[lo:hi], self.r_foot[lo:hi], self.I_world[lo:hi],
                       self.mass[lo:hi], self.t0[lo:hi], self.dt, self.gait_hz, self.duty, self.N)

    def shard(self, rank, world):
        """Contiguous slice of ceil(B/world) robots for ``rank``."""
        per = -(-self.B // world)
        return self.slice(min(rank * per, self.B), min((rank + 1) * per, self.B))


def _rot_zyx(roll, pitch, yaw):
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    R = np.empty(roll.shape + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def host_contact_table(t0, dt, N, gait_hz, duty, phase_offset=PHASE_OFFSET):
    """Batched stance table (B, 4, N) int32 with the reference's arithmetic order (gait.py:26-37).
    Host twin of the device kernel ``cmpc_contact_table``; used to lay out synthetic lever arms."""
    t0 = np.asarray(t0, dtype=np.float64).reshape(-1)
    T = 1 / gait_hz
    t = t0[:, None] + np.arange(N)[None, :] * dt
    t = t + dt / 2
    ph = np.mod(np.asarray(phase_offset)[None, :, None] + t[:, None, :] / T, 1.0)
    return (ph < duty).astype(np.int32)


def random_records(B, N=16, seed=None, gait_hz=3.0, duty=0.6, stress=0.0):
    """Config #3/#5 generator: i.i.d. Go2 trot states and references.

    ``stress`` in [0,1] mixes in large velocity/attitude disturbances (lateral shoves, falls) so
    that friction-pyramid and fz_min constraints become active.
    """
    rng = np.random.default_rng(B if seed is None else seed)
    dt = (1.0 / gait_hz) / N
    t0 = 1e-3 * rng.integers(0, 10000, size=B).astype(np.float64)
    yaw = rng.uniform(-np.pi, np.pi, B)
    roll = np.clip(rng.normal(0, 0.05, B), -0.15, 0.15)
    pitch = np.clip(rng.normal(0, 0.05, B), -0.15, 0.15)
    cmd = np.stack([rng.uniform(-0.8, 0.8, B), rng.uniform(-0.4, 0.4, B), np.zeros(B)], axis=1)
    wz = rng.uniform(-4.0, 4.0, B)
    p = np.stack([rng.uniform(-5, 5, B), rng.uniform(-5, 5, B), 0.27 + rng.normal(0, 0.01, B)], axis=1)
    cy, sy = np.cos(yaw), np.sin(yaw)
    v_des = np.stack([cy * cmd[:, 0] - sy * cmd[:, 1], sy * cmd[:, 0] + cy * cmd[:, 1], np.zeros(B)], axis=1)
    v = v_des + rng.normal(0, 0.1, (B, 3))
    om = np.stack([rng.normal(0, 0.2, B), rng.normal(0, 0.2, B), wz + rng.normal(0, 0.2, B)], axis=1)
    if stress > 0:
        hit = rng.random(B) < stress
        v = v + hit[:, None] * rng.normal(0, 1.0, (B, 3))
        roll = roll + hit * rng.normal(0, 0.2, B)
        pitch = pitch + hit * rng.normal(0, 0.2, B)
    x0 = np.concatenate([p, np.stack([roll, pitch, yaw], axis=1), v, om], axis=1)

    # reference trajectory: com_trajectory.py:47-103 (position target within +-0.1 m of the state)
    pos_des = p.copy()
    pos_des[:, :2] += rng.uniform(-0.1, 0.1, (B, 2))
    pos_des[:, 2] = 0.27
    tv = (np.arange(N) + 1) * dt
    x_ref = np.zeros((B, 12, N))
    x_ref[:, 0:3, :] = pos_des[:, :, None] + v_des[:, :, None] * tv[None, None, :]
    x_ref[:, 5, :] = yaw[:, None] + wz[:, None] * tv[None, :]
    x_ref[:, 6:9, :] = v_des[:, :, None]
    x_ref[:, 11, :] = wz[:, None]

    # inertia: body diag rotated into the world frame
    R = _rot_zyx(roll, pitch, yaw)
    I_world = np.einsum("bij,j,bkj->bik", R, GO2_I_BODY, R)
    mass = np.full(B, GO2_MASS)

    # lever arms: nominal hip position under the body, constant within a stance segment, 0 in swing
    table = host_contact_table(t0, dt, N, gait_hz, duty)                    # (B,4,N)
    hips = np.array([[HIP_X, HIP_Y], [HIP_X, -HIP_Y], [-HIP_X, HIP_Y], [-HIP_X, -HIP_Y]])
    r_foot = np.zeros((B, 4, 3, N))
    # segment id per (b, leg, k): increments at every swing->stance edge
    edge = np.zeros((B, 4, N), dtype=np.int64)
    edge[:, :, 1:] = (table[:, :, 1:] == 1) & (table[:, :, :-1] == 0)
    seg = np.cumsum(edge, axis=2)                                           # 0..2
    noise = rng.normal(0, 0.03, (B, 4, 3, 2))                               # per segment xy jitter
    for leg in range(4):
        hx = cy * hips[leg, 0] - sy * hips[leg, 1]
        hy = sy * hips[leg, 0] + cy * hips[leg, 1]
        jit = np.take_along_axis(noise[:, leg], np.minimum(seg[:, leg], 2)[:, :, None], axis=1)  # (B,N,2)
        st = table[:, leg, :].astype(np.float64)
        r_foot[:, leg, 0, :] = (hx[:, None] + jit[:, :, 0]) * st
        r_foot[:, leg, 1, :] = (hy[:, None] + jit[:, :, 1]) * st
        r_foot[:, leg, 2, :] = (-p[:, 2])[:, None] * st
    return Records(x0, x_ref, r_foot, I_world, mass, t0, dt, gait_hz, duty, N)


GO2_HIP_OFFSET = np.array([[0.1934, 0.0465, 0.0], [0.1934, -0.0465, 0.0], [-0.1934, 0.0465, 0.0], [-0.1934, -0.0465, 0.0]])
GO2_STANCE_OFFSET = np.array([[0.1934, 0.142, 0.0], [0.1934, -0.142, 0.0], [-0.1934, 0.142, 0.0], [-0.1934, -0.142, 0.0]])


def random_cycle_inputs(B, seed=None):
    """Robot states + commands of one MPC cycle as test_MPC.py:173-196 feeds them (host arrays for
    ``CentroidalMPC.cycle_host`` / ``cmpc_cycle_host``): the same distributions as ``random_records`` (SURVEY.md 8d
    config #3), but BEFORE ``ComTraj.generate_traj`` -- 408 bytes per robot instead of the 3.3 KB record."""
    rng = np.random.default_rng(seed)
    so, hip = GO2_STANCE_OFFSET, GO2_HIP_OFFSET
    yaw = rng.uniform(-np.pi, np.pi, B)
    x = np.zeros((B, 12))
    x[:, 0:2] = rng.uniform(-5, 5, (B, 2)); x[:, 2] = 0.27 + rng.normal(0, 0.01, B); x[:, 5] = yaw
    x[:, 3:5] = np.clip(rng.normal(0, 0.05, (B, 2)), -0.15, 0.15)
    cmd = np.stack([rng.uniform(-0.8, 0.8, B), rng.uniform(-0.4, 0.4, B), np.full(B, 0.27), rng.uniform(-4, 4, B)], axis=1)
    c, s_ = np.cos(yaw), np.sin(yaw)
    x[:, 6] = c * cmd[:, 0] - s_ * cmd[:, 1]; x[:, 7] = s_ * cmd[:, 0] + c * cmd[:, 1]
    x[:, 6:9] += rng.normal(0, 0.1, (B, 3))
    x[:, 9:12] = rng.normal(0, 0.2, (B, 3)); x[:, 11] += cmd[:, 3]
    R = _rot_zyx(x[:, 3], x[:, 4], x[:, 5])
    lever = np.zeros((B, 4, 3))
    for leg in range(4):
        lever[:, leg, 0] = c * so[leg, 0] - s_ * so[leg, 1]
        lever[:, leg, 1] = s_ * so[leg, 0] + c * so[leg, 1]
        lever[:, leg, 2] = -x[:, 2]
    lever[:, :, 0:2] += rng.normal(0, 0.03, (B, 4, 2))
    t0 = 1e-3 * rng.integers(0, 10000, B).astype(np.float64)
    return dict(x0=x, R_wb=np.ascontiguousarray(np.swapaxes(R, 1, 2)), lever=lever, cmd=cmd, t0=t0, pos_des=x[:, 0:3].copy(),
                I_world=np.ascontiguousarray(np.einsum("bij,j,bkj->bik", R, GO2_I_BODY, R)), mass=np.full(B, GO2_MASS), hip=hip)


def srb_closed_loop_step(rec, Ad, Bd, gd, u0, mpc_period=0.02, rng=None, noise=0.0):
    """Advance every robot by one MPC period with the first-step forces (stand-in for MuJoCo,
    SURVEY.md section 8f-2): x+ = x + (mpc_period/dt) * (Ad x + Bd[0] u0 + gd - x), then rebuild the
    reference window from the new state.  Used to make warm-start sequences (configs #1/#2)."""
    frac = mpc_period / rec.dt
    x = rec.x0
    xn = np.einsum("bij,bj->bi", Ad, x) + np.einsum("bij,bj->bi", Bd[:, 0], u0) + gd.reshape(-1, 12)
    x_new = x + frac * (xn - x)
    if noise and rng is not None:
        x_new = x_new + rng.normal(0, noise, x_new.shape)
    return x_new


def next_cycle(rec, u_first, mpc_period=0.02):
    """Closed-loop replay (BASELINE configs[1]): advance every robot of ``rec`` by one
    MPC period under its first-step forces ``u_first`` (B,12) with the single-rigid-body model the MPC
    itself uses (com_trajectory.py:234-270, closed-form ZOH), then rebuild the next cycle's record the way
    ``ComTraj.generate_traj`` does (com_trajectory.py:84-106): constant commanded velocity / yaw rate, the
    reference restarted from the new position, contact table at the new time, nominal lever arms.
    Host NumPy: this stands in for MuJoCo + Pinocchio, it is not part of the hot path."""
    B, N, dt = rec.B, rec.N, rec.dt
    x = rec.x0
    yaw_avg = rec.x_ref[:, 5, :].mean(axis=1)
    cy, sy = np.cos(yaw_avg), np.sin(yaw_avg)
    f = u_first.reshape(B, 4, 3)
    F = f.sum(axis=1)
    r0 = rec.r_foot[:, :, :, 0]                                   # (B,4,3) lever arms of step 0
    tau = np.einsum("bij,bj->bi", np.linalg.inv(rec.I_world), np.cross(r0, f).sum(axis=1))
    om = x[:, 9:12]
    rz_om = np.stack([cy * om[:, 0] + sy * om[:, 1], -sy * om[:, 0] + cy * om[:, 1], om[:, 2]], axis=1)
    rz_tau = np.stack([cy * tau[:, 0] + sy * tau[:, 1], -sy * tau[:, 0] + cy * tau[:, 1], tau[:, 2]], axis=1)
    h = dt * dt / 2
    g = np.zeros(3); g[2] = -9.81
    xn = x.copy()
    xn[:, 0:3] += dt * x[:, 6:9] + h * (F / rec.mass[:, None] + g)
    xn[:, 3:6] += dt * rz_om + h * rz_tau
    xn[:, 6:9] += dt * (F / rec.mass[:, None] + g)
    xn[:, 9:12] += dt * tau
    x_new = x + (mpc_period / dt) * (xn - x)
    # next record
    v_des, wz = rec.x_ref[:, 6:9, 0], rec.x_ref[:, 11, 0]
    t0 = rec.t0 + mpc_period
    tv = (np.arange(N) + 1) * dt
    yaw = x_new[:, 5]
    x_ref = np.zeros_like(rec.x_ref)
    pos_des = x_new[:, 0:3].copy()
    pos_des[:, 2] = 0.27
    x_ref[:, 0:3, :] = pos_des[:, :, None] + v_des[:, :, None] * tv[None, None, :]
    x_ref[:, 5, :] = yaw[:, None] + wz[:, None] * tv[None, :]
    x_ref[:, 6:9, :] = v_des[:, :, None]
    x_ref[:, 11, :] = wz[:, None]
    R = _rot_zyx(x_new[:, 3], x_new[:, 4], yaw)
    I_world = np.einsum("bij,j,bkj->bik", R, GO2_I_BODY, R)
    table = host_contact_table(t0, dt, N, rec.gait_hz, rec.duty)
    hips = np.array([[HIP_X, HIP_Y], [HIP_X, -HIP_Y], [-HIP_X, HIP_Y], [-HIP_X, -HIP_Y]])
    c0, s0 = np.cos(yaw), np.sin(yaw)
    r_foot = np.zeros_like(rec.r_foot)
    for leg in range(4):
        st = table[:, leg, :].astype(np.float64)
        r_foot[:, leg, 0, :] = (c0 * hips[leg, 0] - s0 * hips[leg, 1])[:, None] * st
        r_foot[:, leg, 1, :] = (s0 * hips[leg, 0] + c0 * hips[leg, 1])[:, None] * st
        r_foot[:, leg, 2, :] = (-x_new[:, 2])[:, None] * st
    return Records(x_new, x_ref, r_foot, I_world, rec.mass, t0, dt, rec.gait_hz, rec.duty, N)


def retarget(rec, vx_body, vy_body, yaw_rate):
    """New command for every robot of ``rec`` (the piecewise-constant schedule of test_MPC.py:37-47): the reference window is
    rebuilt from the current state the way ``ComTraj.generate_traj`` does (com_trajectory.py:84-103), body-frame velocity
    rotated by the current yaw."""
    N, dt = rec.N, rec.dt
    yaw = rec.x0[:, 5]
    vx_body, vy_body, yaw_rate = (np.broadcast_to(np.asarray(a, dtype=np.float64), yaw.shape) for a in (vx_body, vy_body, yaw_rate))
    c, s_ = np.cos(yaw), np.sin(yaw)
    v = np.stack([c * vx_body - s_ * vy_body, s_ * vx_body + c * vy_body, np.zeros_like(yaw)], axis=1)
    tv = (np.arange(N) + 1) * dt
    x_ref = np.zeros_like(rec.x_ref)
    pos = rec.x0[:, 0:3].copy(); pos[:, 2] = 0.27
    x_ref[:, 0:3, :] = pos[:, :, None] + v[:, :, None] * tv[None, None, :]
    x_ref[:, 5, :] = yaw[:, None] + yaw_rate[:, None] * tv[None, :]
    x_ref[:, 6:9, :] = v[:, :, None]
    x_ref[:, 11, :] = yaw_rate[:, None]
    return Records(rec.x0, x_ref, rec.r_foot, rec.I_world, rec.mass, rec.t0, rec.dt, rec.gait_hz, rec.duty, N)


# ------------------------------------------------------------------------------------------------
# Recorded-states replay format (BASELINE configs[1]): one .npz holds C cycles of B robots,
# every array with a leading cycle axis -- x0 (C,B,12), x_ref (C,B,12,N), r_foot (C,B,4,3,N), I_world (C,B,3,3), mass (C,B),
# t0 (C,B) -- plus dt, gait_hz, duty, N and, optionally, the forces a solver returned (u (C,B,12N)) for later comparison.
# The reference keeps the same per-cycle quantities in Python lists for its plots (test_MPC.py:100-131).
# ------------------------------------------------------------------------------------------------
def save_cycles(path, cycles, u=None):
    first = cycles[0]
    arrs = {k: np.stack([getattr(c, k) for c in cycles]) for k in ("x0", "x_ref", "r_foot", "I_world", "mass", "t0")}
    if u is not None:
        arrs["u"] = np.stack(u)
    np.savez(path, dt=first.dt, gait_hz=first.gait_hz, duty=first.duty, N=first.N, **arrs)


def load_cycles(path):
    """(list of Records, recorded forces (C,B,12N) or None)."""
    z = np.load(path)
    C = z["x0"].shape[0]
    cyc = [Records(z["x0"][c], z["x_ref"][c], z["r_foot"][c], z["I_world"][c], z["mass"][c], z["t0"][c], float(z["dt"]),
                   float(z["gait_hz"]), float(z["duty"]), int(z["N"])) for c in range(C)]
    return cyc, (z["u"] if "u" in z.files else None)


def srb_step_host(x, u_first, x_ref, r_foot, I_world, mass, T, I_body=GO2_I_BODY, stance_offset=None):
    """NumPy twin of ``cmpc_srb_step`` (csrc/cmpc_traj.cuh): the MPC's own single-rigid-body model
    (com_trajectory.py:234-270) held for ``T`` seconds under the first-step forces (exact ZOH, A_c^2 = 0).
    Returns (x_new (B,12), R_world_to_body (B,3,3), I_world (B,3,3), foot_lever (B,4,3)).  Host stand-in for
    MuJoCo + Pinocchio; test / workload-generation use only."""
    if stance_offset is None:
        stance_offset = np.array([[HIP_X, HIP_Y, 0], [HIP_X, -HIP_Y, 0], [-HIP_X, HIP_Y, 0], [-HIP_X, -HIP_Y, 0]])
    B = x.shape[0]
    yaw_avg = x_ref[:, 5, :].mean(axis=1)
    cy, sy = np.cos(yaw_avg), np.sin(yaw_avg)
    f = u_first.reshape(B, 4, 3)
    F = f.sum(axis=1)
    r0 = r_foot[:, :, :, 0]
    al = np.einsum("bij,bj->bi", np.linalg.inv(I_world), np.cross(r0, f).sum(axis=1))
    acc = F / mass[:, None] + np.array([0, 0, -9.81])
    om = x[:, 9:12]
    rz = lambda v: np.stack([cy * v[:, 0] + sy * v[:, 1], -sy * v[:, 0] + cy * v[:, 1], v[:, 2]], axis=1)
    h = T * T / 2
    xn = x.copy()
    xn[:, 0:3] += T * x[:, 6:9] + h * acc
    xn[:, 3:6] += T * rz(om) + h * rz(al)
    xn[:, 6:9] += T * acc
    xn[:, 9:12] += T * al
    R = _rot_zyx(xn[:, 3], xn[:, 4], xn[:, 5])
    I_new = np.einsum("bij,j,bkj->bik", R, np.asarray(I_body), R)
    c0, s0 = np.cos(xn[:, 5]), np.sin(xn[:, 5])
    so = np.asarray(stance_offset)
    lever = np.zeros((B, 4, 3))
    for leg in range(4):
        lever[:, leg, 0] = c0 * so[leg, 0] - s0 * so[leg, 1]
        lever[:, leg, 1] = s0 * so[leg, 0] + c0 * so[leg, 1]
        lever[:, leg, 2] = -xn[:, 2]
    return xn, np.swapaxes(R, 1, 2).copy(), I_new, lever
