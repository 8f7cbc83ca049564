"""Oracle: reference trajectory + lever-arm generator, a NumPy restatement of ``ComTraj.generate_traj``
(reference ``convex_mpc/com_trajectory.py:27-211``) with the gait helpers it calls (``gait.py:21-24``,
``gait.py:40-74``) and the joint-less floating base of ``go2_robot_data.py:224-248`` in closed form.

Pinned by ``tests/golden/reference_traj_vectors.npz``: outputs of the reference's own ``generate_traj``
executed with a stub robot (tests/golden/make_golden_traj.py).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np

PHASE_OFFSET = np.array([0.5, 0.0, 0.0, 0.5])   # gait.py:8


def current_mask(t, hz, duty, offset=PHASE_OFFSET):
    """gait.py:21-24 -> compute_contact_table(t, 0, 1): no half-step offset."""
    period = 1 / hz
    tt = t + np.arange(1) * 0
    tt = tt + 0 / 2
    ph = np.mod(np.asarray(offset)[:, None] + tt[None, :] / period, 1.0)
    return (ph < duty).astype(np.int32)[:, 0]


def generate_traj(x0, R_wb, levers, cmd, t_now, dt, N, hz, duty, hip, pos_des):
    """One robot.  Returns (pos_des_out (3,), x_ref (12,N), r_foot (4,3,N))."""
    x0 = np.asarray(x0, dtype=np.float64)
    pos_des = np.array(pos_des, dtype=np.float64)
    max_err = 0.1                                            # com_trajectory.py:44
    for a in range(2):                                       # :47-56
        if pos_des[a] - x0[a] > max_err:
            pos_des[a] = x0[a] + max_err
        if x0[a] - pos_des[a] > max_err:
            pos_des[a] = x0[a] - max_err
    pos_des[2] = cmd[2]                                      # :58
    yaw, yaw_rate = x0[5], cmd[3]
    Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    vw = Rz @ np.array([cmd[0], cmd[1], 0.0])                # :72
    t_vec = (np.arange(N) + 1) * dt                          # :65
    pos = pos_des.reshape(3, 1) + vw.reshape(3, 1) * t_vec.reshape(1, N)   # :84-86
    vel = np.repeat(vw.reshape(3, 1), N, axis=1)
    rpy = np.zeros((3, N)); rpy[2] = yaw + yaw_rate * t_vec  # :96-98
    om = np.zeros((3, N)); om[2] = yaw_rate                  # :101-103
    x_ref = np.vstack([pos, rpy, vel, om])
    period = 1 / hz
    t_swing, t_stance = (1 - duty) * period, duty * period   # gait.py:18-19
    pred = (t_swing + 0.5 * t_stance) / 2.0                  # gait.py:53-54
    r = np.zeros((4, 3, N))
    nxt = [np.asarray(l, dtype=np.float64).copy() for l in levers]   # :116
    prev = np.array([2, 2, 2, 2])
    v_body = np.asarray(R_wb) @ vw                           # :125-131
    for i in range(N):
        m = current_mask(t_now + i * dt, hz, duty)           # :120
        base = pos[:, i]
        yi = rpy[2, i]
        Rzi = np.array([[np.cos(yi), -np.sin(yi), 0], [np.sin(yi), np.cos(yi), 0], [0, 0, 1]])
        for leg in range(4):
            if m[leg] != prev[leg] and m[leg] == 0:          # take-off (:137-143)
                hip_w = np.array([base[0], base[1], 0]) + Rzi @ hip[leg]          # gait.py:47-48
                nominal = np.array([hip_w[0], hip_w[1], 0.02])
                drift = np.array([v_body[0] * pred, v_body[1] * pred, 0])
                dth = yaw_rate * pred
                rxy = nominal[:2] - base[:2]
                rot = np.array([-dth * rxy[1], dth * rxy[0], 0.0])
                nxt[leg] = (nominal + drift + rot) - base
                r[leg, :, i] = 0.0
            elif m[leg] != prev[leg] and m[leg] == 1:        # touch-down (:145-147)
                r[leg, :, i] = nxt[leg]
            else:                                            # (:149-151)
                r[leg, :, i] = r[leg, :, i - 1]
        prev = m
    return pos_des, x_ref, r
