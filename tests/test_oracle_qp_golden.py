"""Pin the oracle's sparse-QP restatement (rows a5-a12 of SURVEY.md section 8) against the QP data the reference's
own ``CentroidalMPC`` produced (tests/golden/make_golden_qp.py: ``centroidal_mpc.py`` executed verbatim on a
NumPy-backed CasADi container).  Everything is compared bit for bit: these are assignments, products with a single
non-zero term and one 12-term matrix-vector product (``beq_first``, centroidal_mpc.py:259).
"""
import os

import numpy as np
import pytest

from oracle import condensed_qp, sparse_qp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gq():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_qp_vectors.npz"))


def _cases(gq):
    for i in range(int(gq["count"])):
        k = f"qp{i}_"
        yield i, {f: gq[k + f] for f in ("cfg", "x0", "x_ref", "contact", "Ad", "Bd", "gd", "h_diag", "h_nnz", "g", "a_row",
                                         "a_col", "a_val", "a_shape", "lba", "uba", "lbx", "ubx", "warm_is_prev", "solver",
                                         "nvars")}


def test_module_constants(gq):
    """centroidal_mpc.py:12-38."""
    assert np.array_equal(gq["const_Q"], sparse_qp.COST_Q)
    assert np.array_equal(gq["const_R"], sparse_qp.COST_R)
    mu, nx, nu = gq["const_scalars"]
    assert (mu, nx, nu) == (sparse_qp.MU, sparse_qp.NX, sparse_qp.NU)
    assert str(gq["const_solver_name"]) == "osqp"
    eps_abs, eps_rel, max_iter, polish, adaptive, check, interval, scaling, scaled_term, wsp, wsd = gq["const_osqp"]
    assert (eps_abs, eps_rel, max_iter, polish, adaptive, check, interval, scaling, scaled_term, wsp, wsd) == \
        (1e-4, 1e-4, 1000, 0, 1, 10, 25, 5, 1, 1, 1)


def test_at_least_fifty_contact_tables(gq):
    tabs = {c["contact"].tobytes() for _, c in _cases(gq)}
    assert int(gq["count"]) >= 50 and len(tabs) >= 30


def test_sparse_qp_equals_reference_bit_for_bit(gq):
    for i, c in _cases(gq):
        N = int(c["cfg"][0])
        qp = sparse_qp.build(c["Ad"], c["Bd"], c["gd"], c["x0"], c["x_ref"], c["contact"])
        assert qp["H"].shape == (24 * N, 24 * N) == (int(c["nvars"]),) * 2
        assert np.array_equal(qp["H"].diagonal(), c["h_diag"]), i                    # :183-201
        assert qp["H"].nnz == int(c["h_nnz"]) == 24 * N
        assert np.array_equal(qp["g"], c["g"]), i                                    # :248-253
        A = qp["A"].toarray()
        ref = np.zeros(tuple(c["a_shape"]))
        ref[c["a_row"], c["a_col"]] = c["a_val"]
        assert A.shape == ref.shape == (28 * N, 24 * N)
        assert np.array_equal(A, ref), i                                             # :287-303, :324-359
        # the structural pattern too: dynamics rows (identity, -Ad, -Bd blocks are dense in the reference), friction rows
        pat = np.zeros(ref.shape, dtype=bool)
        pat[c["a_row"], c["a_col"]] = True
        mine = sparse_qp.structural_pattern(N)
        assert np.array_equal(pat, mine), i
        for f in ("lba", "uba", "lbx", "ubx"):                                       # :257-282, :122-176
            assert np.array_equal(qp[f], c[f]), (i, f)
        assert c["warm_is_prev"].all()                                               # :92-95, :106-110
        assert str(c["solver"]) == "osqp"


def test_condensed_rows_follow_the_reference_rows(gq):
    """The condensed oracle's constraint rows are the reference's lbx/ubx on the forces followed by its friction rows."""
    for i, c in _cases(gq):
        N = int(c["cfg"][0])
        A, l, u = condensed_qp.constraints(c["contact"])
        assert np.array_equal(l[:12 * N], c["lbx"][12 * N:]) and np.array_equal(u[:12 * N], c["ubx"][12 * N:])
        assert np.array_equal(u[12 * N:], c["uba"][12 * N:]) and np.array_equal(l[12 * N:], c["lba"][12 * N:])
        ref = np.zeros(tuple(c["a_shape"]))
        ref[c["a_row"], c["a_col"]] = c["a_val"]
        assert np.array_equal(A[12 * N:], ref[12 * N:, 12 * N:])
        assert not ref[12 * N:, :12 * N].any()


def test_product_module_constants_equal_reference_execution(gq):
    """The drop-in module re-exports the reference's constants (centroidal_mpc.py:12-38): compare with the values
    frozen from the executed reference module, not with literals."""
    from convex_mpc_b200 import centroidal_mpc as m
    assert np.array_equal(np.diag(m.COST_MATRIX_Q), gq["const_Q"]) and np.array_equal(np.diag(m.COST_MATRIX_R), gq["const_R"])
    assert (m.MU, m.NX, m.NU) == tuple(gq["const_scalars"]) and m.SOLVER_NAME == str(gq["const_solver_name"])
    o = m.OPTS["osqp"]
    mine = [o["eps_abs"], o["eps_rel"], o["max_iter"], float(o["polish"]), float(o["adaptive_rho"]), o["check_termination"],
            o["adaptive_rho_interval"], o["scaling"], float(o["scaled_termination"]), float(m.OPTS["warm_start_primal"]),
            float(m.OPTS["warm_start_dual"])]
    assert np.array_equal(np.array(mine, dtype=np.float64), gq["const_osqp"])
