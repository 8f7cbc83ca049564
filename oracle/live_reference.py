"""The LIVE reference arm: the reference's own ``CentroidalMPC`` (CasADi -> OSQP) executed on records of this
repo's workloads, when the real ``casadi`` wheel and the reference tree are present (SURVEY.md section 8c "L4",
section 8d).

In the authoring image and on the GPU box ``import casadi`` fails (no wheel, no network), so ``available()`` is False
there, the L4 test skips and ``bench.py --impl reference`` falls back to the restated port (``oracle/osqp_port.c``).
On a machine that has both, nothing else is needed: the reference modules are imported from
``$CMPC_REFERENCE_DIR`` (default ``/root/reference/convex_mpc``) with ``go2_robot_data`` (Pinocchio) replaced by an
empty stub -- ``CentroidalMPC`` only uses that module for a type hint -- and a ``ComTraj`` is filled from a record
the way ``generate_traj`` leaves it (com_trajectory.py:84-211), its dynamics by the reference's own
``_continuousDynamics`` / ``_discreteDynamics`` (:221-286).

``casadi_module`` lets the tests run the same plumbing on ``tests/golden/casadi_stub.py`` (with the oracle as the
stub's solver), so the hook is exercised even where the wheel is missing.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import contextlib
import importlib
import io
import os
import sys
import time
import types

import numpy as np

REF_DIR = os.environ.get("CMPC_REFERENCE_DIR", "/root/reference/convex_mpc")


def real_casadi():
    try:
        import casadi
    except Exception:
        return None
    return casadi if hasattr(casadi, "conic") and hasattr(casadi, "__version__") else None


def available(ref_dir=REF_DIR):
    return real_casadi() is not None and os.path.isfile(os.path.join(ref_dir, "centroidal_mpc.py"))


def load(ref_dir=REF_DIR, casadi_module=None):
    """Import the reference's gait / com_trajectory / centroidal_mpc modules (fresh) and return them."""
    if casadi_module is not None:
        sys.modules["casadi"] = casadi_module
    stub = types.ModuleType("go2_robot_data")
    stub.PinGo2Model = type("PinGo2Model", (), {})
    sys.modules["go2_robot_data"] = stub
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    mods = []
    for name in ("gait", "com_trajectory", "centroidal_mpc"):
        sys.modules.pop(name, None)
        mods.append(importlib.import_module(name))
    return tuple(mods)


def traj_from_record(com_trajectory, gait_mod, rec, b):
    """A reference ``ComTraj`` holding robot ``b`` of ``rec`` (records.py layout)."""
    T = com_trajectory.ComTraj.__new__(com_trajectory.ComTraj)
    N = rec.N
    T.N = N
    T.m = float(rec.mass[b])
    T.I_com_world = np.array(rec.I_world[b], dtype=float)
    xr = np.array(rec.x_ref[b], dtype=float)
    T.pos_traj_world, T.rpy_traj_world, T.vel_traj_world, T.omega_traj_world = xr[0:3], xr[3:6], xr[6:9], xr[9:12]
    g = gait_mod.Gait(rec.gait_hz if float(rec.gait_hz) != int(rec.gait_hz) else int(rec.gait_hz), rec.duty)
    T.contact_table = g.compute_contact_table(float(rec.t0[b]), rec.dt, N)
    T.r_fl_foot_world, T.r_fr_foot_world, T.r_rl_foot_world, T.r_rr_foot_world = np.array(rec.r_foot[b], dtype=float)
    T.initial_x_vec = np.array(rec.x0[b], dtype=float).reshape(-1, 1)
    T._continuousDynamics(None)
    T._discreteDynamics(rec.dt)
    return T


class LiveReference:
    def __init__(self, eps=1e-5, ref_dir=REF_DIR, casadi_module=None):
        self.gait, self.com_trajectory, self.cm = load(ref_dir, casadi_module)
        self.cm.OPTS["osqp"]["eps_abs"] = eps        # north star: parity is stated at eps_abs = eps_rel = 1e-5
        self.cm.OPTS["osqp"]["eps_rel"] = eps
        self.mpc = {}

    def solve(self, rec, b, warm=False):
        """One ``solve_QP`` of robot b.  dict(U (12N,) in the order of w[12N:], X (12N,), update_ms, solve_ms, total_s)."""
        T = traj_from_record(self.com_trajectory, self.gait, rec, b)
        t0 = time.perf_counter()
        mpc = self.mpc.get(rec.N)
        if mpc is None:
            with contextlib.redirect_stdout(io.StringIO()):
                mpc = self.mpc[rec.N] = self.cm.CentroidalMPC(None, T)
            t0 = time.perf_counter()                 # construction is outside the reference's own timers too
        if not warm:
            mpc.x_prev = None
        sol = mpc.solve_QP(None, T, False)
        total = time.perf_counter() - t0
        w = np.asarray(sol["x"].full() if hasattr(sol["x"], "full") else sol["x"]).reshape(-1)
        N = rec.N
        stats = mpc.solver.stats()
        return dict(U=w[12 * N:].copy(), X=w[:12 * N].copy(), update_ms=mpc.update_time, solve_ms=mpc.solve_time,
                    total_s=total, status=stats.get("return_status"))
