"""Oracle self-consistency: the condensed QP and the reference's sparse QP are the same problem.

SURVEY.md Appendix B: the condensed optimum lifted to (w, lam_x, lam_a) must satisfy the KKT system
of the sparse QP that ``centroidal_mpc.py`` hands to OSQP.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest

from oracle import condensed_qp, dynamics_ref, exact, gait_ref, sparse_qp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _records():
    name = "cmpc_records_for_tests"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(ROOT, "convex-mpc-unitree-go2_b200", "records.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def make_qp(rec, b):
    ct = gait_ref.contact_table(rec.t0[b], rec.dt, rec.N, rec.gait_hz, rec.duty)
    yaw = dynamics_ref.yaw_average(rec.x_ref[b])
    Ac, Bc, gc = dynamics_ref.continuous_dynamics(rec.mass[b], rec.I_world[b], yaw, rec.r_foot[b])
    Ad, Bd, gd = dynamics_ref.discrete_dynamics_closed(Ac, Bc, gc, rec.dt)
    return ct, Ad, Bd, gd


@pytest.mark.parametrize("stress", [0.0, 1.0])
def test_condensed_equals_sparse(stress):
    rec = _records().random_records(6, seed=11, stress=stress)
    for b in range(rec.B):
        ct, Ad, Bd, gd = make_qp(rec, b)
        cq = condensed_qp.build(Ad, Bd, gd, rec.x0[b], rec.x_ref[b], ct)
        sq = sparse_qp.build(Ad, Bd, gd, rec.x0[b], rec.x_ref[b], ct)
        sol = exact.solve_exact(cq["H"], cq["g"], cq["A"], cq["l"], cq["u"])
        assert sol["ok"]
        k = condensed_qp.kkt_residuals(cq, sol["U"], sol["y"])
        assert k["stat"] < 1e-9 and k["prim"] < 1e-9 and k["dual"] < 1e-9 and k["comp"] < 1e-7
        w, lam_x, lam_a = condensed_qp.lift(cq, sol["U"], sol["y"], Ad=Ad, Bd=Bd, x_ref=rec.x_ref[b])
        # sparse-form KKT: H w + g + lam_x + A' lam_a = 0, equalities hold, bounds hold
        r = sq["H"] @ w + sq["g"] + lam_x + sq["A"].T @ lam_a
        assert np.abs(r).max() < 1e-8
        Aw = sq["A"] @ w
        N = rec.N
        assert np.abs(Aw[:12 * N] - sq["lba"][:12 * N]).max() < 1e-10          # dynamics rows
        assert (Aw[12 * N:] <= sq["uba"][12 * N:] + 1e-9).all()               # friction rows
        assert (w >= sq["lbx"] - 1e-9).all() and (w <= sq["ubx"] + 1e-9).all()
        # objectives agree up to the constant dropped by the sparse form
        js = sparse_qp.objective(sq, w)
        jc = condensed_qp.objective(cq, sol["U"]) + cq["c0"]
        assert abs(js - jc) <= 1e-9 * max(1.0, abs(js))
        X, U = sparse_qp.split_solution(w, N)
        assert U.shape == (12, N) and np.allclose(U.reshape(-1, order="F"), sol["U"])


def test_sparse_structure_counts():
    """Row/column counts and orderings quoted in SURVEY.md section 8 (a7, a8, a11)."""
    rec = _records().random_records(1, seed=3)
    ct, Ad, Bd, gd = make_qp(rec, 0)
    sq = sparse_qp.build(Ad, Bd, gd, rec.x0[0], rec.x_ref[0], ct)
    N = 16
    assert sq["A"].shape == (28 * N, 24 * N) and sq["H"].shape == (24 * N, 24 * N)
    F = sparse_qp.friction_matrix(N).toarray()
    assert F.shape == (256, 384) and (F != 0).sum() == 512
    # row 16k + 4 leg + face, faces (+fx, -fx, +fy, -fy) - mu fz
    k, leg = 5, 2
    r = 16 * k + 4 * leg
    j = 12 * N + 12 * k + 3 * leg
    assert F[r, j] == 1 and F[r + 1, j] == -1 and F[r + 2, j + 1] == 1 and F[r + 3, j + 1] == -1
    assert all(F[r + f, j + 2] == -0.8 for f in range(4))
    swing = ct == 0
    lbx = sq["lbx"][12 * N:].reshape(12, N, order="F")
    ubx = sq["ubx"][12 * N:].reshape(12, N, order="F")
    for leg in range(4):
        assert (lbx[3 * leg:3 * leg + 3][:, swing[leg]] == 0).all()
        assert (ubx[3 * leg:3 * leg + 3][:, swing[leg]] == 0).all()
        assert (lbx[3 * leg + 2, ~swing[leg]] == 10).all()
        assert np.isinf(ubx[3 * leg + 2, ~swing[leg]]).all()
        assert np.isinf(lbx[3 * leg, ~swing[leg]]).all()


def test_exact_optimum_agrees_with_scipy_slsqp():
    """Third-party cross-check of the oracle's optimum (the CasADi -> OSQP binary of centroidal_mpc.py:98 cannot be
    installed here): SciPy's SLSQP, an independent QP/NLP code, solves the same condensed QP over the stance forces
    -- with active friction / fz_min rows -- to the same point (<= 1e-3 N, far inside the 1e-2 N parity budget) and
    the same objective."""
    from scipy.optimize import minimize
    from convex_mpc_b200 import records
    from helpers import oracle_solution
    rec = records.random_records(8, seed=5, stress=0.5)
    n_active_seen = 0
    for b in (4, 5):
        o = oracle_solution(rec, b)
        cq = o["cq"]
        H, g, A, l, u = cq["H"], cq["g"], cq["A"], cq["l"], cq["u"]
        n = H.shape[0]
        free = np.flatnonzero(~((l[:n] == 0) & (u[:n] == 0)))           # swing forces are pinned to zero
        Hf, gf = H[np.ix_(free, free)], g[free]
        rows, rhs = [], []
        for i, j in enumerate(free):
            if np.isfinite(l[j]):
                e = np.zeros(len(free)); e[i] = 1.0
                rows.append(e); rhs.append(l[j])                        # fz >= fz_min
        F, fu = A[n:][:, free], u[n:]
        for r in np.flatnonzero(np.isfinite(fu)):
            rows.append(-F[r]); rhs.append(-fu[r])                      # pyramid face <= 0
        G, hv = np.array(rows), np.array(rhs)
        sc = 1.0 / np.abs(Hf).max()
        x0 = np.zeros(len(free)); x0[2::3] = 40.0
        res = minimize(lambda x: sc * (0.5 * x @ Hf @ x + gf @ x), x0, jac=lambda x: sc * (Hf @ x + gf), method="SLSQP",
                       constraints=[dict(type="ineq", fun=lambda x: G @ x - hv, jac=lambda x: G)],
                       options=dict(ftol=1e-16, maxiter=2000))
        assert res.status == 0
        U = np.zeros(n); U[free] = res.x
        Us = o["sol"]["U"]
        assert np.abs(U - Us).max() < 1e-3
        obj = lambda v: 0.5 * v @ H @ v + g @ v
        assert abs(obj(U) - obj(Us)) < 1e-8 * max(1.0, abs(obj(Us)))
        n_active_seen += int((np.abs(G @ Us[free] - hv) < 1e-7).sum())
    assert n_active_seen > 0            # the cases do exercise active constraints
