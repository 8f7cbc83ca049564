"""Golden vectors for the trajectory / lever-arm generator, made by EXECUTING the
reference's own ``ComTraj.generate_traj`` (com_trajectory.py:27-211) and the gait helpers it calls
(gait.py:21-37, 40-74).

    python tests/golden/make_golden_traj.py          # authoring container only (/root/reference)

Pinocchio is not installed, so ``go2_robot_data.PinGo2Model`` is replaced by a stub that provides exactly
the members ``generate_traj`` reads.  What the stub restates of the real class is only this:
``update_model_simplified(q, dq)`` zeroes the joints and places the floating base at ``q`` (go2_robot_data.py:
224-248), after which ``current_config.base_pos = q[0:3]``, ``current_config.base_vel = dq[0:3]`` and
``R_z`` is the yaw rotation of ``q[5]`` (go2_robot_data.py:212-222); the hip offsets are constants of the
model (go2_robot_data.py:147-161).  Everything else -- reference trajectory, clamp of the position target,
take-off / touch-down state machine, touchdown prediction -- is the reference's code running as written.
Nothing of the reference is copied; only inputs and outputs are frozen into ``reference_traj_vectors.npz``.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/convex_mpc"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_traj_vectors.npz")

HIP = {"FL": np.array([0.1934, 0.0465, 0.0]), "FR": np.array([0.1934, -0.0465, 0.0]),
       "RL": np.array([-0.1934, 0.0465, 0.0]), "RR": np.array([-0.1934, -0.0465, 0.0])}   # placeholder Go2 values


def rot_zyx(roll, pitch, yaw):
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


class StubGo2:
    """Stand-in for PinGo2Model: state set by hand (the 'real robot') or by update_model_simplified (the dummy)."""

    def __init__(self):
        self.current_config = types.SimpleNamespace(base_pos=np.zeros(3), base_vel=np.zeros(3))
        self.data = types.SimpleNamespace(Ig=types.SimpleNamespace(mass=15.0, inertia=np.eye(3)))
        self.R_z = np.eye(3)
        self.R_world_to_body = np.eye(3)
        self.yaw_rate_des_world = []
        self._x = np.zeros(12)
        self._levers = [np.zeros(3)] * 4

    def set_real(self, x, mass, inertia, levers):
        self._x = np.asarray(x, dtype=float).copy()
        yaw = self._x[5]
        self.R_z = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        self.R_world_to_body = rot_zyx(*self._x[3:6]).T
        self.data.Ig.mass, self.data.Ig.inertia = mass, inertia
        self._levers = [np.asarray(l, dtype=float) for l in levers]

    def compute_com_x_vec(self):
        return self._x.reshape(-1, 1).copy()

    def get_foot_lever_world(self):
        return [l.copy() for l in self._levers]

    def get_hip_offset(self, leg):
        return HIP[leg.upper()]

    def update_model_simplified(self, q, dq):
        self.current_config.base_pos = np.array(q[0:3], dtype=float)
        self.current_config.base_vel = np.array(dq[0:3], dtype=float)
        yaw = q[5]
        self.R_z = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])


def load_reference():
    stub = types.ModuleType("go2_robot_data")
    stub.PinGo2Model = StubGo2
    sys.modules["go2_robot_data"] = stub
    sys.path.insert(0, REF)
    import gait            # noqa: E402
    import com_trajectory  # noqa: E402
    return gait, com_trajectory


def main():
    gait_mod, ct_mod = load_reference()
    rng = np.random.default_rng(20261019)
    out = {}
    cfgs = [(3.0, 0.6, 16), (3.0, 0.6, 32), (2.5, 0.5, 16)]
    ncase = 0
    for hz, duty, N in cfgs:
        g = gait_mod.Gait(hz, duty)
        dt = g.gait_period / N
        for rep in range(24):
            go2 = StubGo2()
            yaw = rng.uniform(-np.pi, np.pi)
            x = np.concatenate([[rng.uniform(-5, 5), rng.uniform(-5, 5), 0.27 + rng.normal(0, 0.01)],
                                [rng.normal(0, 0.05), rng.normal(0, 0.05), yaw],
                                rng.normal(0, 0.4, 3), rng.normal(0, 0.3, 3)])
            A = rng.normal(size=(3, 3)) * 0.02
            inertia = np.diag([0.11, 0.33, 0.38]) + A @ A.T
            Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
            levers = [Rz @ (HIP[l] + np.array([0, 0.09 * np.sign(HIP[l][1]), 0])) + np.array([0, 0, -x[2]]) + rng.normal(0, 0.02, 3)
                      for l in ("FL", "FR", "RL", "RR")]
            go2.set_real(x, 15.02 + rng.normal(0, 0.3), inertia, levers)
            traj = ct_mod.ComTraj(go2)
            # the position target the object carries from earlier cycles: sometimes outside the clamp window
            traj.pos_des_world = x[0:3] + rng.uniform(-0.25, 0.25, 3)
            pos_des_in = traj.pos_des_world.copy()
            t_now = float(1e-3 * rng.integers(0, 10000)) if rep % 3 else float(rng.uniform(0, 10))
            cmd = np.array([rng.uniform(-0.8, 0.8), rng.uniform(-0.4, 0.4), 0.27, rng.uniform(-4, 4)])
            traj.generate_traj(go2, g, t_now, cmd[0], cmd[1], cmd[2], cmd[3], dt)
            assert traj.N == N, (traj.N, N)
            k = f"tr{ncase}_"
            out[k + "cfg"] = np.array([hz, duty, N, dt])
            out[k + "x0"] = x
            out[k + "R_wb"] = go2.R_world_to_body
            out[k + "levers"] = np.stack(levers)
            out[k + "mass"] = np.array(go2.data.Ig.mass)
            out[k + "inertia"] = inertia
            out[k + "pos_des_in"] = pos_des_in
            out[k + "cmd"] = cmd
            out[k + "t_now"] = np.array(t_now)
            out[k + "pos_des_out"] = traj.pos_des_world.copy()
            out[k + "x_ref"] = traj.compute_x_ref_vec()
            out[k + "contact"] = traj.contact_table.astype(np.int32)
            out[k + "r_foot"] = np.stack([traj.r_fl_foot_world, traj.r_fr_foot_world, traj.r_rl_foot_world, traj.r_rr_foot_world])
            out[k + "Ad"], out[k + "Bd"], out[k + "gd"] = traj.Ad, traj.Bd, traj.gd
            ncase += 1
    out["count"] = np.array(ncase)
    out["hip"] = np.stack([HIP[l] for l in ("FL", "FR", "RL", "RR")])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, ncase, "cases", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
