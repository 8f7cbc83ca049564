"""CPU check of the *kernel source* (csrc/cmpc_core.cuh compiled for the host with a one-thread CTA,
tests/_emul) against the oracle.  Races and launch plumbing are only visible on the GPU (-m gpu), but
indexing, recursions and the solver control flow are identical code."""
import numpy as np
import pytest

from convex_mpc_b200 import records
from helpers import Emul, force_error, oracle_inputs, oracle_solution
from oracle import condensed_qp, gait_ref, sparse_qp


@pytest.fixture(scope="module")
def emul():
    return Emul()


def test_contact_table_bit_exact_vs_golden(emul, golden):
    for ci in range(int(golden["ct_count"])):
        hz, duty, N, dt = golden[f"ct{ci}_cfg"]
        N = int(N)
        if N > 48:
            continue
        mask = emul.contact_table(golden[f"ct{ci}_t0"], float(dt), N, float(hz), float(duty))
        got = gait_ref.unpack_mask(mask, N)
        assert np.array_equal(got, golden[f"ct{ci}_table"]), ci


def test_contact_table_bit_exact_random(emul):
    rng = np.random.default_rng(5)
    t0 = np.concatenate([1e-3 * rng.integers(0, 100000, 20000), rng.uniform(0, 50, 20000)])
    dt = (1 / 3) / 16
    mask = emul.contact_table(t0, dt, 16, 3, 0.6)
    got = gait_ref.unpack_mask(mask, 16)
    for i in range(0, t0.size, 97):
        assert np.array_equal(got[i], gait_ref.contact_table(t0[i], dt, 16, 3, 0.6))


def test_dense_build_matches_oracle(emul):
    rec = records.random_records(4, seed=21)
    H, g = emul.build(rec)
    for b in range(rec.B):
        ct, Ad, Bd, gd = oracle_inputs(rec, b)
        cq = condensed_qp.build(Ad, Bd, gd, rec.x0[b], rec.x_ref[b], np.ones((4, rec.N), dtype=int))
        scale = np.abs(cq["H"]).max()
        assert np.abs(H[b] - cq["H"]).max() <= 1e-12 * scale
        assert np.abs(g[b] - cq["g"]).max() <= 1e-12 * np.abs(cq["g"]).max()
    # drop-in path: the same data handed over as (Ad, Bd, gd)
    AB = [oracle_inputs(rec, b) for b in range(rec.B)]
    H2, g2 = emul.build(rec, Ad=np.stack([a[1] for a in AB]), Bd=np.stack([a[2] for a in AB]),
                        gd=np.stack([a[3].reshape(12) for a in AB]))
    assert np.abs(H2 - H).max() <= 1e-12 * np.abs(H).max()
    assert np.abs(g2 - g).max() <= 1e-11 * np.abs(g).max()


@pytest.mark.parametrize("impl", ["generic", "fast"])
@pytest.mark.parametrize("stress", [0.0, 0.5, 1.0])
def test_forces_match_exact_optimum(emul, stress, impl):
    rec = records.random_records(24, seed=100 + int(10 * stress), stress=stress)
    r = emul.solve(rec) if impl == "generic" else emul.solve_fast(rec)
    assert (r["status"] == 1).all()
    for b in range(rec.B):
        o = oracle_solution(rec, b)
        assert o["sol"]["ok"]
        assert np.array_equal(gait_ref.unpack_mask(r["mask"][b], rec.N), o["ct"])
        err, rel = force_error(r["u"][b], o["sol"]["U"])
        assert rel < 1e-3, (b, err)            # 1000x inside the 1e-2 N / 1e-3 tolerance
        assert np.abs(r["y"][b] - o["sol"]["y"]).max() < 1e-7
        # lifted solution and cost in the reference's own (sparse) form
        sq = sparse_qp.build(o["Ad"], o["Bd"], o["gd"], rec.x0[b], rec.x_ref[b], o["ct"])
        w = np.concatenate([r["X"][b], r["u"][b]])
        assert abs(sparse_qp.objective(sq, w) - r["stats"][b, 2]) <= 1e-9 * max(1, abs(r["stats"][b, 2]))
        Aw = sq["A"] @ w
        assert np.abs(Aw[:12 * rec.N] - sq["lba"][:12 * rec.N]).max() < 1e-10
        lam_x = np.concatenate([np.zeros(12 * rec.N), r["y"][b][:12 * rec.N]])
        lam_a = np.concatenate([r["nu"][b], r["y"][b][12 * rec.N:]])
        assert np.abs(sq["H"] @ w + sq["g"] + lam_x + sq["A"].T @ lam_a).max() < 1e-8


@pytest.mark.parametrize("impl", ["generic", "fast"])
def test_admm_mode_is_osqp_like(emul, impl):
    """mode 0 = plain ADMM with OSQP's termination rule: residuals below eps, forces only roughly right
    (that is how OSQP itself behaves on this flat QP)."""
    rec = records.random_records(8, seed=31, stress=0.3)
    solve = emul.solve if impl == "generic" else emul.solve_fast
    r = solve(rec, mode=0, eps_abs=1e-6, eps_rel=1e-6, max_iter=4000)
    assert (r["status"] == 1).all() and (r["iters"] > 0).all()
    for b in range(rec.B):
        o = oracle_solution(rec, b)
        assert np.abs(r["u"][b] - o["sol"]["U"]).max() < 0.2
        assert r["stats"][b, 0] < 1e-3 and r["stats"][b, 1] < 1e-4
    # polish on top of ADMM recovers the exact optimum
    r2 = solve(rec, mode=0, polish=1, eps_abs=1e-4, eps_rel=1e-4)
    for b in range(rec.B):
        o = oracle_solution(rec, b)
        assert force_error(r2["u"][b], o["sol"]["U"])[1] < 1e-3


@pytest.mark.parametrize("impl", ["generic", "fast"])
def test_warm_start_and_edge_masks(emul, impl):
    rec = records.random_records(6, seed=41, stress=0.5)
    solve = emul.solve if impl == "generic" else emul.solve_fast
    cold = solve(rec)
    warm = solve(rec, warm=1, state=(cold["u"].copy(), cold["y"].copy(), cold["rho"].copy()))
    assert np.abs(warm["u"] - cold["u"]).max() < 1e-7
    assert (warm["stats"][:, 6] <= np.maximum(cold["stats"][:, 6], 1)).all()
    N = rec.N
    W = 1
    # all swing: zero forces, multipliers close stationarity; all stance: 192 free variables
    for bits, nfree in ((0, 0), (2 ** 64 - 1, 192)):
        mask = np.full((rec.B, W), bits, dtype=np.uint64)
        r = solve(rec, mask=mask)
        assert (r["status"] == 1).all() and (r["stats"][:, 3] == nfree).all()
        ct = np.full((4, N), 1 if bits else 0)
        for b in range(2):
            o = oracle_solution(rec, b, contact=ct)
            assert force_error(r["u"][b], o["sol"]["U"])[1] < 1e-3
            assert np.abs(r["y"][b] - o["sol"]["y"]).max() < 1e-7
    # a bound on stance foot-steps that is too small is reported, never silently truncated
    r = solve(rec, nfmax=8)
    assert (r["status"] == -20).all()


def test_fast_path_equals_generic_path_all_horizons(emul):
    """The closed-form fast path (cmpc_fast.cuh) and the recursion-based generic path (cmpc_core.cuh)
    are independent derivations of the same QP: forces, duals, states, co-states and cost agree."""
    for N, B in ((16, 16), (32, 4), (48, 2)):
        rec = records.random_records(B, N=N, seed=300 + N, stress=0.4)
        a, b = emul.solve(rec), emul.solve_fast(rec)
        assert (a["status"] == 1).all() and (b["status"] == 1).all()
        assert np.abs(a["u"] - b["u"]).max() < 1e-6
        assert np.abs(a["y"] - b["y"]).max() < 1e-8
        assert np.abs(a["X"] - b["X"]).max() < 1e-9
        assert np.abs(a["nu"] - b["nu"]).max() < 1e-7
        assert np.abs(a["stats"][:, 2] - b["stats"][:, 2]).max() < 1e-8 * max(1.0, np.abs(a["stats"][:, 2]).max())
        assert np.array_equal(a["stats"][:, 3], b["stats"][:, 3])


@pytest.mark.parametrize("N", [16, 32])
def test_riccati_prepass_finishes_exactly_the_unconstrained_robots(emul, N):
    """csrc/cmpc_riccati.cuh: the pre-pass must finish the robots whose optimum has no active inequality --
    with the condensed path's outputs -- and leave every other robot untouched for the condensed kernel."""
    rec = records.random_records(24, N=N, seed=5, stress=0.2)
    a = emul.riccati(rec)
    b = emul.solve_fast(rec)
    done = a["done"].astype(bool)
    assert np.array_equal(done, b["stats"][:, 7] == 0)           # same set as PATH_UNCONSTRAINED
    assert done.any() and (~done).any()
    for key, tol in (("u", 1e-8), ("y", 1e-8), ("X", 1e-9), ("nu", 1e-8)):
        assert np.abs(a[key][done] - b[key][done]).max() < tol, key
    assert (a["status"][done] == 1).all() and a["stats"][done, 1].max() < 1e-10     # r_dual certificate
    assert np.abs(a["stats"][done, 2] - b["stats"][done, 2]).max() < 1e-9           # objective
    assert np.abs(a["u"][~done]).max() == 0.0 and (a["status"][~done] == 0).all()   # nothing written
    for bi in np.flatnonzero(done)[:3]:
        o = oracle_solution(rec, bi)
        assert force_error(a["u"][bi], o["sol"]["U"])[1] < 1e-3


def test_traj_generator_matches_reference_golden(emul):
    """csrc/cmpc_traj.cuh (host emulation) against the reference's own generate_traj outputs."""
    from helpers import golden_traj_batches
    for g in golden_traj_batches():
        pd, xr, rf = emul.generate_traj(g["N"], g["x0"], g["R_wb"], g["levers"], g["cmd"], g["t_now"], g["dt"], g["hz"],
                                        g["duty"], g["hip"], g["pos_des_in"])
        assert np.array_equal(pd, g["pos_des_out"])                            # clamp: exact
        assert np.array_equal(rf == 0.0, g["r_foot"] == 0.0)                   # take-off / touch-down pattern: exact
        assert np.abs(xr - g["x_ref"]).max() <= 1e-13 * max(1.0, np.abs(g["x_ref"]).max())
        assert np.abs(rf - g["r_foot"]).max() <= 1e-14


def test_srb_step_matches_numpy_twin(emul):
    """csrc/cmpc_traj.cuh srb_step_one (host emulation) against records.srb_step_host."""
    import ctypes
    from helpers import _p
    rec = records.random_records(64, seed=9, stress=0.3)
    rng = np.random.default_rng(3)
    u = rng.normal(0, 30, (rec.B, 12 * rec.N)); u[:, 2:12:3] = np.abs(u[:, 2:12:3]) + 20
    so = np.array([[0.19, 0.14, 0], [0.19, -0.14, 0], [-0.19, 0.14, 0], [-0.19, -0.14, 0]], dtype=np.float64)
    ib = np.array([0.11, 0.33, 0.38])
    xo = np.zeros((rec.B, 12)); Rw = np.zeros((rec.B, 3, 3)); Io = np.zeros((rec.B, 3, 3)); lv = np.zeros((rec.B, 4, 3))
    emul.lib.emul_srb_step(rec.N, rec.B, _p(rec.x0), _p(u), _p(rec.x_ref), _p(rec.r_foot), _p(rec.I_world), _p(rec.mass),
                           ctypes.c_double(0.02), _p(ib), _p(so), _p(xo), _p(Rw), _p(Io), _p(lv))
    x2, R2, I2, l2 = records.srb_step_host(rec.x0, u[:, :12], rec.x_ref, rec.r_foot, rec.I_world, rec.mass, 0.02, ib, so)
    for a, b in ((xo, x2), (Rw, R2), (Io, I2), (lv, l2)):
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max())


def test_tighter_stance_bound_keeps_the_working_set_capacity(emul):
    """A tighter ``max_stance`` only shrinks the workspace: heavily constrained robots at N = 32 (working sets of up
    to 78 rows under the general bound 4N) must take the same route and reach the same forces under the periodic-gait
    bound 4(floor(0.6 N) + 1) = 80 (csrc/cmpc_fast.cuh: kcap_fast)."""
    rec = records.random_records(8, N=32, seed=902, stress=1.0)
    a, b = emul.solve_fast(rec, nfmax=4 * rec.N), emul.solve_fast(rec, nfmax=80)
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["stats"][:, 7], b["stats"][:, 7])
    assert np.abs(a["u"] - b["u"]).max() < 1e-9
    # the case is only meaningful if an active-set robot holds more than 64 rows (the capacity 80 foot-steps alone give)
    rows = []
    for i in range(rec.B):
        ct = gait_ref.unpack_mask(a["mask"][i], rec.N)
        y = a["y"][i]
        fz_rows = sum(abs(y[12 * k + 3 * leg + 2]) > 1e-9 for k in range(rec.N) for leg in range(4) if ct[leg, k])
        rows.append(int(fz_rows + (np.abs(y[12 * rec.N:]) > 1e-9).sum()))
    assert max(r for r, path in zip(rows, a["stats"][:, 7]) if path == 1) > 64, rows


def test_leg_jacobian_source_against_oracle_and_finite_differences():
    """traj::leg_jacobian (the device source, compiled for the host) against oracle/leg_kin.py (joint axis x lever, a
    different derivation) and against central differences of the oracle's link-by-link forward kinematics -- the
    analytic counterpart of compute_3x3_foot_Jacobian_world (go2_robot_data.py:286-300)."""
    from oracle import leg_kin
    from convex_mpc_b200 import records
    rng = np.random.default_rng(3)
    B = 64
    q = np.stack([rng.uniform(-0.8, 0.8, (B, 4)), rng.uniform(-0.5, 2.0, (B, 4)), rng.uniform(-2.6, -0.9, (B, 4))], axis=2).reshape(B, 12)
    R = records._rot_zyx(rng.normal(0, 0.2, B), rng.normal(0, 0.2, B), rng.uniform(-np.pi, np.pi, B))
    R_wb = np.ascontiguousarray(np.swapaxes(R, 1, 2))
    J, pb = Emul().leg_jacobian(q, R_wb, leg_kin.GO2_LINKS)
    h = 1e-6
    for b in range(B):
        for leg in range(4):
            q3 = q[b, 3 * leg:3 * leg + 3]
            side = leg_kin.SIDE[leg]
            assert np.abs(pb[b, leg] - leg_kin.foot_pos_body(q3, side)).max() < 1e-15
            Jo = leg_kin.jacobian_world(q3, R_wb[b], side)
            assert np.abs(J[b, leg] - Jo).max() < 1e-15
            fd = np.stack([(leg_kin.foot_pos_body(q3 + h * e, side) - leg_kin.foot_pos_body(q3 - h * e, side)) / (2 * h)
                           for e in np.eye(3)], axis=1)
            assert np.abs(J[b, leg] - R_wb[b].T @ fd).max() < 1e-9
    # nominal stance (hip 0, thigh 0.9, calf -1.8): the foot is under the thigh joint, 0.0955 m out from the hip
    _, p0 = Emul().leg_jacobian(np.tile([0.0, 0.9, -1.8], 4)[None], np.eye(3)[None], leg_kin.GO2_LINKS)
    assert np.abs(p0[0, :, 0]).max() < 1e-15 and np.allclose(p0[0, :, 1], [0.0955, -0.0955, 0.0955, -0.0955])
    assert np.allclose(p0[0, :, 2], -2 * 0.213 * np.cos(0.9))
