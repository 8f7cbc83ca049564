"""L4: forces against the reference's own CasADi -> OSQP solve on identical records.

The real ``casadi`` wheel is not installable in the authoring image nor on the GPU box, so the L4 comparison skips
there; the plumbing of the hook (``oracle/live_reference.py``: reference modules imported, ``ComTraj`` filled from a
record, ``CentroidalMPC.solve_QP`` called, forces extracted) is still exercised wherever the reference tree exists, on
the NumPy-backed CasADi container of tests/golden with the oracle standing in for OSQP.
"""
import os
import sys

import numpy as np
import pytest

from oracle import live_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_TREE = os.path.isfile(os.path.join(live_reference.REF_DIR, "centroidal_mpc.py"))


def _records():
    from convex_mpc_b200 import records
    return records


@pytest.mark.skipif(not HAVE_TREE, reason="reference tree not present (GPU box)")
def test_live_hook_plumbing_on_the_casadi_container():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import casadi_stub
    from oracle import exact, sparse_qp
    import helpers

    def hook(qp, opts, kw):
        # the oracle stands in for OSQP: solve the condensed twin of the QP the reference just assembled
        hook.calls += 1
        hook.opts = opts
        sol = helpers.oracle_solution(hook.rec, hook.b)
        N = hook.rec.N
        from oracle import condensed_qp
        w = condensed_qp.lift(sol["cq"], sol["sol"]["U"])
        # ... and make sure the reference's g / bounds describe that same problem
        sq = sparse_qp.build(sol["Ad"], sol["Bd"], sol["gd"], hook.rec.x0[hook.b], hook.rec.x_ref[hook.b], sol["ct"])
        assert np.array_equal(kw["g"].a.reshape(-1), sq["g"]) and np.array_equal(kw["lbx"].a.reshape(-1), sq["lbx"])
        assert np.array_equal(kw["uba"].a.reshape(-1), sq["uba"])
        hook.U = sol["sol"]["U"]
        return {"x": casadi_stub.DM(w), "lam_x": casadi_stub.DM.zeros(24 * N), "lam_a": casadi_stub.DM.zeros(28 * N)}

    hook.calls = 0
    casadi_stub.solve_hook = hook
    try:
        live = live_reference.LiveReference(eps=1e-5, casadi_module=casadi_stub)
        rec = _records().random_records(3, seed=5, stress=0.5)
        for b in range(rec.B):
            hook.rec, hook.b = rec, b
            r = live.solve(rec, b)
            assert np.array_equal(r["U"], hook.U) and r["X"].shape == (12 * rec.N,)
            assert r["solve_ms"] >= 0 and r["update_ms"] >= 0
        assert hook.calls == 3 and hook.opts["osqp"]["eps_abs"] == 1e-5 and hook.opts["osqp"]["eps_rel"] == 1e-5
    finally:
        casadi_stub.solve_hook = None
        for name in ("casadi", "gait", "com_trajectory", "centroidal_mpc", "go2_robot_data"):
            sys.modules.pop(name, None)


@pytest.mark.skipif(not live_reference.available(), reason="real casadi wheel and/or reference tree not present")
def test_oracle_optimum_vs_live_casadi_osqp():
    """Forces of the oracle's exact optimum vs the reference's OSQP at eps 1e-5: OSQP stops at its termination
    test, so it is held to its own tolerance, the oracle to the KKT certificate."""
    import helpers
    live = live_reference.LiveReference(eps=1e-5)
    rec = _records().random_records(8, seed=9)
    for b in range(rec.B):
        r = live.solve(rec, b)
        sol = helpers.oracle_solution(rec, b)
        # objective of the reference's answer can not beat the exact optimum by more than OSQP's tolerance allows
        from oracle import condensed_qp
        j_ref = condensed_qp.objective(sol["cq"], r["U"])
        j_opt = condensed_qp.objective(sol["cq"], sol["sol"]["U"])
        assert j_ref >= j_opt - 1e-3 * max(1.0, abs(j_opt))
        assert np.isfinite(r["U"]).all()
