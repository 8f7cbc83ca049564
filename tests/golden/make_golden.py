"""Generate golden vectors by EXECUTING the reference's own NumPy/SciPy code.

Run in the authoring container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules import Pinocchio through ``go2_robot_data``; a stub module stands in for it so
that ``gait.Gait.compute_contact_table`` (gait.py:26-37) and
``com_trajectory.ComTraj._continuousDynamics/_discreteDynamics/compute_x_ref_vec``
(com_trajectory.py:15-25,221-286) run *verbatim*.  Nothing from the reference is copied into the repo;
only its numerical outputs on seeded inputs are frozen into ``reference_vectors.npz``.

The QP solve (CasADi -> OSQP, centroidal_mpc.py:98) cannot be executed here (casadi is not
installed), so no golden forces exist: that row stays "parity unpinned".
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/convex_mpc"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")


def load_reference():
    stub = types.ModuleType("go2_robot_data")
    stub.PinGo2Model = type("PinGo2Model", (), {})
    sys.modules["go2_robot_data"] = stub
    sys.path.insert(0, REF)
    import gait            # noqa: E402
    import com_trajectory  # noqa: E402
    return gait, com_trajectory


def main():
    gait, com_trajectory = load_reference()
    rng = np.random.default_rng(20261018)
    out = {}

    # ---- contact tables: (hz, duty, N) x many t0 ------------------------------------------
    cfgs = [(3, 0.6, 16), (3.0, 0.6, 32), (3, 0.6, 48), (2.5, 0.5, 16), (4, 0.75, 16), (3, 0.6, 1)]
    for ci, (hz, duty, N) in enumerate(cfgs):
        g = gait.Gait(hz, duty)
        dt = g.gait_period / N if N > 1 else 0.0
        t0 = np.concatenate([
            1e-3 * rng.integers(0, 10000, size=400),          # the sim's 1 kHz time stamps
            rng.uniform(0, 10, size=400),                     # arbitrary times
            np.arange(0, 64) * (1.0 / hz) / 64,               # phase boundaries
            np.array([0.0, 1.0 / 3.0, 0.2, 0.1 + 0.2, 1e-9, 123.456]),
        ])
        tabs = np.stack([g.compute_contact_table(float(t), dt, N) for t in t0])
        out[f"ct{ci}_cfg"] = np.array([hz, duty, N, dt], dtype=np.float64)
        out[f"ct{ci}_t0"] = t0
        out[f"ct{ci}_table"] = tabs.astype(np.int32)
    out["ct_count"] = np.array(len(cfgs))

    # ---- dynamics: random but plausible SRB inputs ----------------------------------------
    cases = []
    for ci, N in enumerate([16, 16, 16, 32, 48, 16, 16, 16]):
        T = com_trajectory.ComTraj.__new__(com_trajectory.ComTraj)
        dt = (1.0 / 3.0) / N
        yaw0 = rng.uniform(-np.pi, np.pi)
        wz = rng.uniform(-4, 4)
        tv = (np.arange(N) + 1) * dt
        T.N = N
        T.m = 15.02 + rng.normal(0, 0.5)
        A = rng.normal(size=(3, 3)) * 0.02
        T.I_com_world = np.diag([0.11, 0.33, 0.38]) + A @ A.T
        T.pos_traj_world = rng.normal(size=(3, 1)) + rng.normal(size=(3, 1)) * tv[None, :]
        T.vel_traj_world = np.repeat(rng.normal(size=(3, 1)), N, axis=1)
        T.rpy_traj_world = np.zeros((3, N))
        T.rpy_traj_world[2, :] = yaw0 + wz * tv
        T.omega_traj_world = np.zeros((3, N))
        T.omega_traj_world[2, :] = wz
        feet = rng.normal(size=(4, 3, N)) * 0.2
        feet[:, :, rng.integers(0, N, size=3)] = 0.0          # swing steps have zero lever arms
        T.r_fl_foot_world, T.r_fr_foot_world, T.r_rl_foot_world, T.r_rr_foot_world = feet
        T._continuousDynamics(None)
        T._discreteDynamics(dt)
        out[f"dyn{ci}_in_scalar"] = np.array([N, dt, T.m])
        out[f"dyn{ci}_I"] = T.I_com_world
        out[f"dyn{ci}_rfoot"] = feet
        out[f"dyn{ci}_xref"] = T.compute_x_ref_vec()
        out[f"dyn{ci}_Ac"] = T.Ac
        out[f"dyn{ci}_Bc"] = T.Bc
        out[f"dyn{ci}_Ad"] = T.Ad
        out[f"dyn{ci}_Bd"] = T.Bd
        out[f"dyn{ci}_gd"] = T.gd
        cases.append(ci)
    out["dyn_count"] = np.array(len(cases))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
