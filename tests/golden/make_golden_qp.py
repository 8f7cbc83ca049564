"""Golden vectors for the QP DATA of the hot path, made by EXECUTING the
reference's own ``CentroidalMPC`` (``convex_mpc/centroidal_mpc.py``): ``__init__`` (:41-67), ``solve_QP``
(:69-120), ``_compute_bounds`` (:122-176), ``_build_sparse_matrix`` (:178-233), ``_update_sparse_matrix``
(:235-285), ``_assemble_A_matrix`` (:287-303), ``_create_dynamics_function`` (:305-321) and
``_precompute_friction_matrix`` (:324-359) -- all verbatim, with

* ``casadi``          replaced by ``tests/golden/casadi_stub.py`` (a NumPy container; the wheel is not installable),
* ``go2_robot_data``  replaced by an empty stub (Pinocchio is not installable; the class is only a type hint here),
* ``traj``            the reference's own ``ComTraj`` object, its trajectory members set by hand and its dynamics
                      filled by the reference's ``_continuousDynamics`` / ``_discreteDynamics``.

    python tests/golden/make_golden_qp.py          # authoring container only (/root/reference)

What is frozen per case (``reference_qp_vectors.npz``): the inputs (``x0, x_ref, contact, Ad, Bd, gd``) and what
``solve_QP`` hands to ``ca.conic``: ``h`` (diagonal), ``g``, ``a`` (non-zeros as triplets), ``lba, uba, lbx, ubx``;
plus the warm-start hand-over of the second call (``x0, lam_x0, lam_a0`` are the previous ``sol``).  The module
constants (``COST_MATRIX_Q/R, MU, NX, NU, OPTS, SOLVER_NAME``) are frozen once.  Nothing of the reference is copied.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/convex_mpc"
OUT = os.path.join(HERE, "reference_qp_vectors.npz")


def load_reference():
    sys.path.insert(0, HERE)
    import casadi_stub
    sys.modules["casadi"] = casadi_stub
    stub = types.ModuleType("go2_robot_data")
    stub.PinGo2Model = type("PinGo2Model", (), {})
    sys.modules["go2_robot_data"] = stub
    sys.path.insert(0, REF)
    import gait             # noqa: E402
    import com_trajectory   # noqa: E402
    import centroidal_mpc   # noqa: E402
    return casadi_stub, gait, com_trajectory, centroidal_mpc


def make_traj(com_trajectory, gait_mod, rng, N, hz, duty, t0):
    """A ComTraj in the state generate_traj leaves it in (com_trajectory.py:84-211), members set by hand."""
    T = com_trajectory.ComTraj.__new__(com_trajectory.ComTraj)
    g = gait_mod.Gait(hz, duty)
    dt = g.gait_period / N
    tv = (np.arange(N) + 1) * dt
    yaw0, wz = rng.uniform(-np.pi, np.pi), rng.uniform(-4, 4)
    v = rng.normal(size=(3, 1)) * np.array([[0.5], [0.3], [0.0]])
    p0 = rng.normal(size=(3, 1)) * np.array([[2.0], [2.0], [0.0]]) + np.array([[0], [0], [0.27]])
    T.N = N
    T.m = 15.02 + rng.normal(0, 0.3)
    A = rng.normal(size=(3, 3)) * 0.02
    T.I_com_world = np.diag([0.11, 0.33, 0.38]) + A @ A.T
    T.pos_traj_world = p0 + v * tv[None, :]
    T.vel_traj_world = np.repeat(v, N, axis=1)
    T.rpy_traj_world = np.zeros((3, N))
    T.rpy_traj_world[2, :] = yaw0 + wz * tv
    T.omega_traj_world = np.zeros((3, N))
    T.omega_traj_world[2, :] = wz
    T.contact_table = g.compute_contact_table(float(t0), dt, N)          # com_trajectory.py:106, gait.py:26-37
    feet = np.zeros((4, 3, N))
    hips = np.array([[0.19, 0.14], [0.19, -0.14], [-0.19, 0.14], [-0.19, -0.14]])
    c, s = np.cos(yaw0), np.sin(yaw0)
    for leg in range(4):
        off = rng.normal(0, 0.03, 2)
        feet[leg, 0] = (c * hips[leg, 0] - s * hips[leg, 1] + off[0]) * T.contact_table[leg]
        feet[leg, 1] = (s * hips[leg, 0] + c * hips[leg, 1] + off[1]) * T.contact_table[leg]
        feet[leg, 2] = -0.27 * T.contact_table[leg]
    T.r_fl_foot_world, T.r_fr_foot_world, T.r_rl_foot_world, T.r_rr_foot_world = feet
    x0 = np.concatenate([p0[:, 0] + rng.normal(0, 0.02, 3), rng.normal(0, 0.05, 2), [yaw0],
                         v[:, 0] + rng.normal(0, 0.1, 3), rng.normal(0, 0.2, 2), [wz + rng.normal(0, 0.2)]])
    T.initial_x_vec = x0.reshape(-1, 1)                                   # com_trajectory.py:37 (12,1)
    T._continuousDynamics(None)                                           # com_trajectory.py:221-270
    T._discreteDynamics(dt)                                               # com_trajectory.py:272-286
    return T, dt, feet


def triplets(dm):
    r, c = np.nonzero(dm.nz)
    return r.astype(np.int32), c.astype(np.int32), dm.a[r, c]


def main():
    ca, gait_mod, com_trajectory, cm = load_reference()
    rng = np.random.default_rng(20261019)
    out = {}
    out["const_Q"] = np.diag(cm.COST_MATRIX_Q).astype(np.float64)
    out["const_R"] = np.diag(cm.COST_MATRIX_R).astype(np.float64)
    out["const_scalars"] = np.array([cm.MU, cm.NX, cm.NU], dtype=np.float64)
    o = cm.OPTS["osqp"]
    out["const_osqp"] = np.array([o["eps_abs"], o["eps_rel"], o["max_iter"], float(o["polish"]), float(o["adaptive_rho"]),
                                  o["check_termination"], o["adaptive_rho_interval"], o["scaling"],
                                  float(o["scaled_termination"]), float(cm.OPTS["warm_start_primal"]),
                                  float(cm.OPTS["warm_start_dual"])])
    out["const_solver_name"] = np.array(cm.SOLVER_NAME)

    cfgs = [(16, 3, 0.6)] * 40 + [(16, 2.5, 0.5)] * 6 + [(16, 4, 0.75)] * 6 + [(32, 3, 0.6)] * 4 + [(48, 3, 0.6)] * 2 + [(4, 3, 0.6)] * 4
    n = 0
    for ci, (N, hz, duty) in enumerate(cfgs):
        t0 = 1e-3 * rng.integers(0, 10000) if ci % 3 else rng.uniform(0, 10)
        T, dt, feet = make_traj(com_trajectory, gait_mod, rng, N, hz, duty, t0)
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            mpc = cm.CentroidalMPC(None, T)                               # centroidal_mpc.py:41-67
        # first call: cold (no x_prev) ; the stub solver returns recognisable vectors
        nv, na = 24 * N, 28 * N
        fake = {"x": ca.DM(np.arange(nv, dtype=float)), "lam_x": ca.DM(-np.arange(nv, dtype=float)),
                "lam_a": ca.DM(0.5 * np.arange(na, dtype=float))}
        ca.solve_hook = lambda qp, opts, kw, _f=fake: _f
        sol = mpc.solve_QP(None, T, False)                                # centroidal_mpc.py:69-120
        call1 = dict(ca.last_call)
        assert "x0" not in call1 and sol is fake
        mpc.solve_QP(None, T, False)
        call2 = dict(ca.last_call)
        k = f"qp{n}_"
        out[k + "cfg"] = np.array([N, hz, duty, dt, t0], dtype=np.float64)
        out[k + "x0"] = T.initial_x_vec.reshape(-1)
        out[k + "x_ref"] = T.compute_x_ref_vec()
        out[k + "contact"] = np.asarray(T.contact_table, dtype=np.int32)
        out[k + "Ad"], out[k + "Bd"], out[k + "gd"] = T.Ad, T.Bd, np.asarray(T.gd).reshape(-1)
        h = call1["h"]
        assert (h.a == np.diag(np.diag(h.a))).all()
        out[k + "h_diag"] = np.diag(h.a).copy()
        out[k + "h_nnz"] = np.array(h.nnz())
        out[k + "g"] = call1["g"].a.reshape(-1)
        ar, ac, av = triplets(call1["a"])
        out[k + "a_row"], out[k + "a_col"], out[k + "a_val"] = ar, ac, av
        out[k + "a_shape"] = np.array(call1["a"].size())
        for f in ("lba", "uba", "lbx", "ubx"):
            out[k + f] = call1[f].a.reshape(-1)
        # warm-start hand-over of the second call (centroidal_mpc.py:92-95,106-110)
        out[k + "warm_is_prev"] = np.array([call2["x0"] is fake["x"], call2["lam_x0"] is fake["lam_x"],
                                            call2["lam_a0"] is fake["lam_a"]])
        out[k + "solver"] = np.array(call1["_solver"])
        out[k + "nvars"] = np.array(mpc.nvars)
        n += 1
    out["count"] = np.array(n)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", n, "cases")


if __name__ == "__main__":
    main()
