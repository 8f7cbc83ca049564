"""The OSQP restatement (oracle/osqp_ref.py) and its C port (oracle/osqp_port.c, the CPU baseline) on
the reference's sparse QP: same iterates, OSQP's own termination test satisfied when re-evaluated
independently, objective consistent with the exact optimum."""
import numpy as np

from convex_mpc_b200 import records
from helpers import oracle_solution
from oracle import cpu_port, osqp_ref, sparse_qp


def _sparse(rec, b):
    o = oracle_solution(rec, b)
    sq = sparse_qp.build(o["Ad"], o["Bd"], o["gd"], rec.x0[b], rec.x_ref[b], o["ct"])
    return o, sq, sparse_qp.as_osqp_form(sq)


def test_port_matches_numpy_restatement_and_terminates_properly():
    rec = records.random_records(3, seed=4, stress=0.5)
    eps = 1e-5
    r = cpu_port.solve_batch(rec, cpu_port.default_opts(eps_abs=eps, eps_rel=eps, max_iter=4000))
    for b in range(rec.B):
        o, sq, (P, q, A, l, u) = _sparse(rec, b)
        ref = osqp_ref.solve(P, q, A, l, u, eps_abs=eps, eps_rel=eps, max_iter=4000)
        assert ref["status"] == "solved" and r["status"][b] == 1
        assert r["iters"][b] == ref["iters"] and r["nfac"][b] == ref["nfac"]
        assert np.abs(r["w"][b] - ref["x"]).max() < 1e-7
        assert np.abs(r["y"][b] - ref["y"]).max() < 1e-7
        assert abs(r["obj"][b] - sparse_qp.objective(sq, r["w"][b])) < 1e-9 * max(1.0, abs(r["obj"][b]))
        # near-optimality in the objective, while forces may still be N away (flat QP)
        w_star = np.concatenate([np.zeros(12 * rec.N), o["sol"]["U"]])
        from oracle import condensed_qp
        w_star[:12 * rec.N] = condensed_qp.rollout(o["cq"], o["sol"]["U"])
        J_star = sparse_qp.objective(sq, w_star)
        assert abs(r["obj"][b] - J_star) < 2e-2 * max(1.0, abs(J_star))
        assert np.abs(r["w"][b][12 * rec.N:] - o["sol"]["U"]).max() < 10.0


def test_port_warm_start_and_threads():
    rec = records.random_records(6, seed=8)
    opts = cpu_port.default_opts(eps_abs=1e-4, eps_rel=1e-4)          # the reference's own OPTS
    cold = cpu_port.solve_batch(rec, opts, nthreads=1)
    par = cpu_port.solve_batch(rec, opts, nthreads=3)
    assert np.array_equal(cold["iters"], par["iters"]) and np.array_equal(cold["w"], par["w"])
    warm = cpu_port.solve_batch(rec, opts, warm=True, state=(cold["w"].copy(), cold["y"].copy(), cold["rho"].copy()))
    assert (warm["status"] == 1).all() and warm["iters"].mean() < 0.5 * cold["iters"].mean()
