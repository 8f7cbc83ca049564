"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI / the drop-in class, against the
oracle on the same seeded inputs.  Levels follow SURVEY.md section 8(c):
  L0 bit-exact contact tables;  L1 A_d/B_d/g_d, H, g (<= 1e-12 rel);
  L2 forces vs the exact optimum (1e-2 N abs + 1e-3 rel, the north-star tolerance);
  L3 OSQP termination test at eps = 1e-5 on the reference's sparse QP.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from convex_mpc_b200 import records  # noqa: E402
from helpers import force_error, oracle_inputs, oracle_solution  # noqa: E402
from oracle import condensed_qp, gait_ref, sparse_qp  # noqa: E402

TOL_ABS, TOL_REL = 1e-2, 1e-3      # BASELINE.json north_star: 1e-3 relative / 1e-2 N absolute


@pytest.fixture(scope="module")
def mod():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from convex_mpc_b200 import centroidal_mpc
    return centroidal_mpc


def make_mpc(mod, rec, **kw):
    traj = mod.BatchedComTraj.from_records(rec, device="cuda:0")
    return mod.CentroidalMPC(None, traj, verbose=False, **kw), traj


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------
def test_contact_table_bit_exact(mod, golden):
    from convex_mpc_b200 import _lib
    lib = _lib.load()
    for ci in range(int(golden["ct_count"])):
        hz, duty, N, dt = golden[f"ct{ci}_cfg"]
        N = int(N)
        t0 = golden[f"ct{ci}_t0"]
        h = ctypes.c_void_p()
        _lib.check(lib.cmpc_create(N, len(t0), 0, ctypes.byref(h)))
        W = (4 * N + 63) // 64
        mask = torch.zeros(len(t0), W, dtype=torch.int64, device="cuda")
        t0_d = dev(t0)          # keep device inputs referenced until the kernel has run
        _lib.check(lib.cmpc_contact_table(h, len(t0), t0_d.data_ptr(), float(dt), float(hz), float(duty),
                                          _lib.darr([0.5, 0, 0, 0.5]), mask.data_ptr(), None))
        torch.cuda.synchronize()
        got = gait_ref.unpack_mask(mask.cpu().numpy().view(np.uint64), N)
        assert np.array_equal(got, golden[f"ct{ci}_table"]), f"contact table config {ci}"
        lib.cmpc_destroy(h)


def test_contact_table_bit_exact_large_random(mod):
    from convex_mpc_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(77)
    B = 1 << 18
    t0 = np.concatenate([1e-3 * rng.integers(0, 10 ** 6, B // 2), rng.uniform(0, 1000, B // 2)])
    dt = (1 / 3) / 16
    h = ctypes.c_void_p()
    _lib.check(lib.cmpc_create(16, B, 0, ctypes.byref(h)))
    mask = torch.zeros(B, 1, dtype=torch.int64, device="cuda")
    t0_d = dev(t0)
    _lib.check(lib.cmpc_contact_table(h, B, t0_d.data_ptr(), dt, 3.0, 0.6, _lib.darr([0.5, 0, 0, 0.5]),
                                      mask.data_ptr(), None))
    torch.cuda.synchronize()
    got = mask.cpu().numpy().view(np.uint64)
    want = gait_ref.pack_mask(records.host_contact_table(t0, dt, 16, 3.0, 0.6))
    assert np.array_equal(got, want)
    # spot-check the vectorised host twin against the scalar oracle too
    for i in range(0, B, 4099):
        assert np.array_equal(gait_ref.unpack_mask(got[i], 16), gait_ref.contact_table(t0[i], dt, 16, 3.0, 0.6))
    lib.cmpc_destroy(h)


def test_dynamics_match_reference_golden(mod, golden):
    from convex_mpc_b200 import _lib
    lib = _lib.load()
    for ci in range(int(golden["dyn_count"])):
        N, dt, m = golden[f"dyn{ci}_in_scalar"]
        N = int(N)
        h = ctypes.c_void_p()
        _lib.check(lib.cmpc_create(N, 1, 0, ctypes.byref(h)))
        Ad = torch.zeros(1, 12, 12, dtype=torch.float64, device="cuda")
        Bd = torch.zeros(1, N, 12, 12, dtype=torch.float64, device="cuda")
        gd = torch.zeros(1, 12, dtype=torch.float64, device="cuda")
        xr, rf, Iw, ms = (dev(golden[f"dyn{ci}_xref"]), dev(golden[f"dyn{ci}_rfoot"]), dev(golden[f"dyn{ci}_I"]),
                          dev(np.array([m])))
        _lib.check(lib.cmpc_dynamics(h, 1, xr.data_ptr(), rf.data_ptr(), Iw.data_ptr(), ms.data_ptr(), float(dt),
                                     Ad.data_ptr(), Bd.data_ptr(), gd.data_ptr(), None))
        torch.cuda.synchronize()
        refB = golden[f"dyn{ci}_Bd"]
        assert np.abs(Ad.cpu().numpy()[0] - golden[f"dyn{ci}_Ad"]).max() <= 1e-15
        assert np.abs(Bd.cpu().numpy()[0] - refB).max() <= 1e-12 * max(1.0, np.abs(refB).max())
        assert np.abs(gd.cpu().numpy()[0] - golden[f"dyn{ci}_gd"].reshape(12)).max() <= 1e-15
        lib.cmpc_destroy(h)


def test_build_H_g_match_oracle(mod):
    from convex_mpc_b200 import _lib
    lib = _lib.load()
    rec = records.random_records(6, seed=21)
    B, N = rec.B, rec.N
    n = 12 * N
    h = ctypes.c_void_p()
    _lib.check(lib.cmpc_create(N, B, 0, ctypes.byref(h)))
    H = torch.zeros(B, n, n, dtype=torch.float64, device="cuda")
    g = torch.zeros(B, n, dtype=torch.float64, device="cuda")
    x0, xr, rf, I, m = dev(rec.x0), dev(rec.x_ref), dev(rec.r_foot), dev(rec.I_world), dev(rec.mass)
    _lib.check(lib.cmpc_build(h, B, None, None, None, x0.data_ptr(), xr.data_ptr(), rf.data_ptr(), I.data_ptr(),
                              m.data_ptr(), rec.dt, H.data_ptr(), g.data_ptr(), None))
    AB = [oracle_inputs(rec, b) for b in range(B)]
    Ad = dev(np.stack([a[1] for a in AB])); Bd = dev(np.stack([a[2] for a in AB]))
    gd = dev(np.stack([a[3].reshape(12) for a in AB]))
    H2 = torch.zeros_like(H); g2 = torch.zeros_like(g)
    _lib.check(lib.cmpc_build(h, B, Ad.data_ptr(), Bd.data_ptr(), gd.data_ptr(), x0.data_ptr(), xr.data_ptr(),
                              None, None, None, rec.dt, H2.data_ptr(), g2.data_ptr(), None))
    torch.cuda.synchronize()
    Hn, gn, H2n, g2n = H.cpu().numpy(), g.cpu().numpy(), H2.cpu().numpy(), g2.cpu().numpy()
    for b in range(B):
        cq = condensed_qp.build(AB[b][1], AB[b][2], AB[b][3], rec.x0[b], rec.x_ref[b], np.ones((4, N), dtype=int))
        for Hx, gx in ((Hn[b], gn[b]), (H2n[b], g2n[b])):
            assert np.abs(Hx - cq["H"]).max() <= 1e-12 * np.abs(cq["H"]).max()
            assert np.abs(gx - cq["g"]).max() <= 1e-11 * np.abs(cq["g"]).max()
            assert np.array_equal(Hx, Hx.T)
    lib.cmpc_destroy(h)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("stress,B", [(0.0, 96), (0.5, 96), (1.0, 64)])
def test_forces_match_exact_optimum(mod, stress, B):
    rec = records.random_records(B, seed=500 + int(10 * stress), stress=stress)
    mpc, traj = make_mpc(mod, rec)
    sol = mpc.solve_QP(None, traj)
    u = sol["u"].cpu().numpy()                       # (B,12,N)
    w = sol["x"].full()                              # (B,24N)
    status = sol["status"].cpu().numpy()
    stats = sol["stats"].cpu().numpy()
    assert (status == 1).all()
    N = rec.N
    worst = 0.0
    for b in range(B):
        o = oracle_solution(rec, b)
        assert o["sol"]["ok"]
        U = u[b].reshape(-1, order="F")              # back to w[12N:] order
        assert np.array_equal(U, w[b, 12 * N:])
        d = np.abs(U - o["sol"]["U"])
        assert (d <= TOL_ABS + TOL_REL * np.abs(o["sol"]["U"])).all(), (b, d.max())
        worst = max(worst, d.max())
        # L3: the lifted point passes OSQP's own termination test at eps = 1e-5 on the reference's QP
        sq = sparse_qp.build(o["Ad"], o["Bd"], o["gd"], rec.x0[b], rec.x_ref[b], o["ct"])
        lam_x = sol["lam_x"].full()[b]
        lam_a = sol["lam_a"].full()[b]
        eps = 1e-5
        P, q, A, l, uu = sparse_qp.as_osqp_form(sq)
        y = np.concatenate([lam_x, lam_a])
        Ax = A @ w[b]
        z = np.clip(Ax, l, uu)
        r_p = np.abs(Ax - z).max()
        r_d = np.abs(P @ w[b] + q + A.T @ y).max()
        assert r_p <= eps + eps * max(np.abs(Ax).max(), np.abs(z).max())
        assert r_d <= eps + eps * max(np.abs(P @ w[b]).max(), np.abs(A.T @ y).max(), np.abs(q).max())
        assert abs(sparse_qp.objective(sq, w[b]) - stats[b, 2]) <= 1e-8 * max(1.0, abs(stats[b, 2]))
    assert worst < 1e-6        # in practice the active-set path is exact to ~1e-9 N


@pytest.mark.parametrize("version", [1, 2, 3])
def test_riccati_prepass_same_optimum(mod, version):
    """cmpc_set_prepass(1): warp-per-robot Riccati sweep + device work-list for the rest.  Same statuses, paths
    and forces as the condensed kernel alone, forces within tolerance of the oracle's exact optimum."""
    rec = records.random_records(2048, seed=611, stress=0.3)
    a, traj = make_mpc(mod, rec, prepass=version, max_stance=40)
    b, _ = make_mpc(mod, rec, max_stance=40, prepass=0)
    sa, sb = a.solve_QP(None, traj), b.solve_QP(None, traj)
    ua, ub = sa["u"].cpu().numpy(), sb["u"].cpu().numpy()
    assert (sa["status"].cpu().numpy() == 1).all() and (sb["status"].cpu().numpy() == 1).all()
    sta, stb = sa["stats"].cpu().numpy(), sb["stats"].cpu().numpy()
    assert np.array_equal(np.where(sta[:, 7] == 4, 0, sta[:, 7]), stb[:, 7])   # path 4 = finished by the pre-pass = unconstrained
    assert 0.2 < (sta[:, 7] == 4).mean() < 0.9                   # both kinds present
    assert np.abs(ua - ub).max() < 1e-7
    assert np.abs(sa["x"].full() - sb["x"].full()).max() < 1e-7
    assert np.abs(sa["lam_a"].full() - sb["lam_a"].full()).max() < 1e-6
    assert np.abs(sta[:, 0] - stb[:, 0]).max() < 1e-9 and sta[:, 1].max() < 1e-6
    assert sta[sta[:, 7] == 4, 1].max() < 1e-10                  # certificate of the Riccati-finished robots
    for bi in range(0, 2048, 256):
        o = oracle_solution(rec, bi)
        assert force_error(ua[bi].reshape(-1, order="F"), o["sol"]["U"])[1] < 1e-3
    # warm start of the second cycle goes through the same two kernels
    sa2 = a.solve_QP(None, traj)
    exact = stb[:, 7] < 2                                        # (ADMM-fallback robots are only eps-accurate)
    assert np.abs(sa2["u"].cpu().numpy() - ub)[exact].max() < 1e-7


def test_traj_generator_matches_reference_golden_and_feeds_the_mpc(mod):
    """cmpc_generate_traj against outputs of the reference's own ComTraj.generate_traj
    (tests/golden/make_golden_traj.py), then straight into solve_QP: the A_d, B_d, g_d the reference derived from
    its trajectory give the same forces through the Ad/Bd route."""
    from convex_mpc_b200 import com_trajectory as ct
    from helpers import golden_traj_batches
    for g in golden_traj_batches():
        st = ct.RobotState(dev(g["x0"]), dev(g["R_wb"]), dev(g["levers"]), dev(g["mass"]), dev(g["inertia"]))
        traj = ct.ComTraj(st, hip_offset=g["hip"], device="cuda:0")
        traj.pos_des_world.copy_(dev(g["pos_des_in"]))
        gait = ct.Gait(g["hz"], g["duty"])
        traj.generate_traj(st, gait, dev(g["t_now"]), dev(g["cmd"][:, 0]), dev(g["cmd"][:, 1]), dev(g["cmd"][:, 2]),
                           dev(g["cmd"][:, 3]), g["dt"])
        assert traj.N == g["N"]
        xr, rf = traj.compute_x_ref_vec().cpu().numpy(), traj.r_foot.cpu().numpy()
        assert np.array_equal(traj.pos_des_world.cpu().numpy(), g["pos_des_out"])
        assert np.array_equal(rf == 0.0, g["r_foot"] == 0.0)
        assert np.abs(xr - g["x_ref"]).max() <= 1e-13 * max(1.0, np.abs(g["x_ref"]).max())
        assert np.abs(rf - g["r_foot"]).max() <= 1e-14
        # second call reuses the carried position target like the reference object does
        traj.generate_traj(st, gait, dev(g["t_now"]), dev(g["cmd"][:, 0]), dev(g["cmd"][:, 1]), dev(g["cmd"][:, 2]),
                           dev(g["cmd"][:, 3]), g["dt"])
        assert np.array_equal(traj.pos_des_world.cpu().numpy(), g["pos_des_out"])
        mpc = mod.CentroidalMPC(None, traj, verbose=False)
        sol = mpc.solve_QP(None, traj)
        assert np.array_equal(gait_ref.unpack_mask(mpc._mask.cpu().numpy().view(np.uint64), g["N"]), g["contact"])
        ab = mod.BatchedComTraj(g["N"], dev(g["x0"]), dev(g["x_ref"]), g["dt"], Ad=dev(g["Ad"]), Bd=dev(g["Bd"]),
                                gd=dev(g["gd"]), contact_table=dev(g["contact"]))
        sol2 = mod.CentroidalMPC(None, ab, verbose=False).solve_QP(None, ab)
        ok = (sol["status"].cpu().numpy() == 1) & (sol2["status"].cpu().numpy() == 1)
        ok &= ~np.isin(sol["stats"].cpu().numpy()[:, 7], (2, 3)) & ~np.isin(sol2["stats"].cpu().numpy()[:, 7], (2, 3))   # exact paths only
        assert ok.mean() > 0.7
        assert np.abs(sol["u"].cpu().numpy()[ok] - sol2["u"].cpu().numpy()[ok]).max() < 1e-5


def test_device_resident_closed_loop(mod):
    """generate_traj -> solve_QP -> srb_step on the device for 12 cycles: the SRB step against
    its NumPy twin, the generated trajectory against the oracle restatement of the reference's generate_traj, and
    the forces of the last cycle against the oracle's exact optimum."""
    from convex_mpc_b200 import com_trajectory as ct
    from oracle import traj_ref
    B, N, HZ, DUTY, T = 96, 16, 3.0, 0.6, 0.02
    rng = np.random.default_rng(77)
    gait = ct.Gait(HZ, DUTY)
    dt = gait.gait_period / N
    hip = np.array([[0.1934, 0.0465, 0], [0.1934, -0.0465, 0], [-0.1934, 0.0465, 0], [-0.1934, -0.0465, 0]])
    so = np.array([[0.1934, 0.142, 0], [0.1934, -0.142, 0], [-0.1934, 0.142, 0], [-0.1934, -0.142, 0]])
    yaw = rng.uniform(-np.pi, np.pi, B)
    x = np.zeros((B, 12)); x[:, 0:2] = rng.uniform(-2, 2, (B, 2)); x[:, 2] = 0.27; x[:, 5] = yaw
    R = records._rot_zyx(x[:, 3], x[:, 4], x[:, 5])
    lever = np.zeros((B, 4, 3))
    for leg in range(4):
        lever[:, leg, 0] = np.cos(yaw) * so[leg, 0] - np.sin(yaw) * so[leg, 1]
        lever[:, leg, 1] = np.sin(yaw) * so[leg, 0] + np.cos(yaw) * so[leg, 1]
        lever[:, leg, 2] = -0.27
    state = ct.RobotState(dev(x), dev(np.swapaxes(R, 1, 2).copy()), dev(lever), dev(np.full(B, records.GO2_MASS)),
                          dev(np.einsum("bij,j,bkj->bik", R, records.GO2_I_BODY, R)))
    cmd = np.stack([rng.uniform(-0.8, 0.8, B), rng.uniform(-0.4, 0.4, B), np.full(B, 0.27), rng.uniform(-2, 2, B)], axis=1)
    cmd_d = [dev(cmd[:, i]) for i in range(4)]
    traj = ct.ComTraj(state, hip_offset=hip, device="cuda:0")
    traj.generate_traj(state, gait, 0.0, *cmd_d, dt)
    mpc = mod.CentroidalMPC(None, traj, verbose=False)
    for c in range(12):
        pos_des_before = traj.pos_des_world.cpu().numpy().copy()
        traj.generate_traj(state, gait, c * T, *cmd_d, dt)
        sol = mpc.solve_QP(None, traj)
        assert (sol["status"].cpu().numpy() == 1).all(), c
        xs, Rs, ls = state.x.cpu().numpy(), state.R_world_to_body.cpu().numpy(), state.foot_lever_world.cpu().numpy()
        xr, rf = traj.compute_x_ref_vec().cpu().numpy(), traj.r_foot.cpu().numpy()
        for b in (0, 17, 95):       # the generated trajectory = the reference's generate_traj (oracle restatement)
            pd, xr_o, rf_o = traj_ref.generate_traj(xs[b], Rs[b], ls[b], cmd[b], c * T, dt, N, HZ, DUTY, hip, pos_des_before[b])
            assert np.abs(xr[b] - xr_o).max() < 1e-12 and np.abs(rf[b] - rf_o).max() < 1e-13
            assert np.array_equal(rf[b] == 0, rf_o == 0)
        u = mpc._u.cpu().numpy().reshape(B, 12 * N)
        x2, R2, I2, l2 = records.srb_step_host(xs, u[:, :12], xr, rf, state.inertia.cpu().numpy(), state.mass.cpu().numpy(),
                                               T, records.GO2_I_BODY, so)
        state = ct.srb_step(state, traj, mpc._u, T, records.GO2_I_BODY, so)
        for a, b_ in ((state.x, x2), (state.R_world_to_body, R2), (state.inertia, I2), (state.foot_lever_world, l2)):
            assert np.abs(a.cpu().numpy() - b_).max() < 1e-11
    rec = records.Records(xs, xr, rf, I2 * 0 + traj.I_com_world.cpu().numpy(), state.mass.cpu().numpy(),
                          np.full(B, 11 * T), dt, HZ, DUTY, N)
    for b in (3, 50):
        o = oracle_solution(rec, b)
        assert force_error(u[b], o["sol"]["U"])[1] < 1e-3
    assert 0.2 < state.x[:, 2].mean().item() < 0.35           # still walking, not fallen through the floor


@pytest.mark.parametrize("N", [32, 48])
def test_riccati_prepass_longer_horizons(mod, N):
    """The lock-step pre-pass with fewer robots per CTA (20 KB / 28 KB of shared memory per robot at N = 32 / 48)
    against the condensed kernel alone and the oracle."""
    B = 2048
    rec = records.random_records(B, N=N, seed=700 + N, stress=0.2)
    ms = 4 * (int(np.floor(rec.duty * N)) + 1)
    a, traj = make_mpc(mod, rec, max_stance=ms, prepass=3)       # the lock-step sweep of round 1 (the default is now 4)
    b, _ = make_mpc(mod, rec, max_stance=ms, prepass=0)
    sa, sb = a.solve_QP(None, traj), b.solve_QP(None, traj)
    sta, stb = sa["stats"].cpu().numpy(), sb["stats"].cpu().numpy()
    assert (sa["status"].cpu().numpy() == 1).all() and (sb["status"].cpu().numpy() == 1).all()
    assert np.array_equal(np.where(sta[:, 7] == 4, 0, sta[:, 7]), stb[:, 7])
    assert (sta[:, 7] == 4).mean() > 0.2
    exact = ~np.isin(stb[:, 7], (2, 3))
    ua, ub = sa["u"].cpu().numpy(), sb["u"].cpu().numpy()
    assert np.abs(ua - ub)[exact].max() < 1e-6
    assert np.abs(sa["lam_a"].full() - sb["lam_a"].full())[exact].max() < 1e-5
    for bi in (0, 777):
        o = oracle_solution(rec, bi)
        assert force_error(ua[bi].reshape(-1, order="F"), o["sol"]["U"])[1] < 1e-3


@pytest.mark.parametrize("version", [1, 2, 3])
def test_riccati_prepass_irregular_contact_schedules(mod, version):
    """Contact tables the trot never produces -- a flight phase (no leg in stance for two steps), single-leg and
    three-leg stance steps, all four legs -- drive the pre-pass through its run-time-m stage code (m = 0, 3, 9, 12)
    and the all-stance case; results against the condensed kernel alone and the oracle."""
    B, N = 2048, 16
    rec = records.random_records(B, seed=313, stress=0.0)
    rng = np.random.default_rng(7)
    ct = np.ones((B, 4, N), dtype=np.int32)
    ct[: B // 2] = (rng.random((B // 2, 4, N)) < 0.7).astype(np.int32)      # arbitrary patterns: 0..4 legs per step
    ct[: B // 2, :, 5:7] = 0                                                 # flight phase
    ct[: B // 2, 1:, 9] = 0; ct[: B // 2, 0, 9] = 1                          # one leg only
    rf = rec.r_foot.copy()
    hips = np.array([[0.1934, 0.142], [0.1934, -0.142], [-0.1934, 0.142], [-0.1934, -0.142]])
    for leg in range(4):                                                     # lever arms wherever the table says stance
        rf[:, leg, 0, :] = hips[leg, 0]; rf[:, leg, 1, :] = hips[leg, 1]; rf[:, leg, 2, :] = -0.27
    rec = records.Records(rec.x0, rec.x_ref, rf * ct[:, :, None, :], rec.I_world, rec.mass, rec.t0, rec.dt, rec.gait_hz, rec.duty, N)
    sols = []
    for pp in (version, 0):
        traj = mod.BatchedComTraj.from_records(rec, device="cuda:0")
        traj.contact_table = dev(ct)
        mpc = mod.CentroidalMPC(None, traj, verbose=False, prepass=pp)
        sols.append(mpc.solve_QP(None, traj))
    sa, sb = sols
    sta, stb = sa["stats"].cpu().numpy(), sb["stats"].cpu().numpy()
    assert np.array_equal(sa["status"].cpu().numpy(), sb["status"].cpu().numpy())
    ok = (sb["status"].cpu().numpy() == 1) & ~np.isin(stb[:, 7], (2, 3))
    assert ok.mean() > 0.8
    assert np.array_equal(np.where(sta[:, 7] == 4, 0, sta[:, 7])[ok], stb[ok, 7])
    assert (sta[: B // 2, 7] == 4).sum() > 20 and (sta[B // 2:, 7] == 4).sum() > 20     # both halves reach the pre-pass
    ua, ub = sa["u"].cpu().numpy(), sb["u"].cpu().numpy()
    assert np.abs(ua - ub)[ok].max() < 1e-6
    assert np.abs(sa["x"].full() - sb["x"].full())[ok].max() < 1e-6
    assert np.abs(sa["lam_a"].full() - sb["lam_a"].full())[ok].max() < 1e-5
    done = np.flatnonzero(ok & (sta[:, 7] == 4))
    for bi in (done[0], done[len(done) // 2], done[-1]):
        o = oracle_solution(rec, bi, contact=ct[bi])
        assert force_error(ua[bi].reshape(-1, order="F"), o["sol"]["U"])[1] < 1e-3


def test_enqueue_and_cuda_graph_replay_equal_solve_QP(mod):
    """CentroidalMPC.enqueue (no host synchronisation) gives solve_QP's results, eagerly and replayed from a CUDA
    graph (the capture path tools/closed_loop.py uses)."""
    rec = records.random_records(512, seed=91, stress=0.3)
    a, traj = make_mpc(mod, rec, max_stance=40)
    ref = a.solve_QP(None, traj)["u"].cpu().numpy().copy()
    b, _ = make_mpc(mod, rec, max_stance=40)
    b.enqueue(traj)
    torch.cuda.synchronize()
    assert np.array_equal(b._u.view(512, 16, 12).transpose(1, 2).cpu().numpy(), ref)
    c, _ = make_mpc(mod, rec, max_stance=40)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        c.enqueue(traj)                  # warm-up: allocations happen outside the capture
        c.reset()
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            c.enqueue(traj)
        c._u.zero_()
        c._rho.zero_()
        g.replay()
        side.synchronize()
    assert (c._status.cpu().numpy() == 1).all()
    exact = ~np.isin(c._stats.cpu().numpy()[:, 7], (2, 3))
    assert np.abs(c._u.view(512, 16, 12).transpose(1, 2).cpu().numpy() - ref)[exact].max() < 1e-7


def test_stance_torque_mapping(mod):
    """cmpc_stance_torque against leg_controller.py:100-101 + test_MPC.py:227 restated in NumPy;
    the current-time gait mask against the reference-pinned oracle (gait.py:21-24)."""
    from convex_mpc_b200 import com_trajectory as ct
    from oracle import traj_ref
    rng = np.random.default_rng(12)
    B, N = 4096, 16
    J = rng.normal(0, 0.2, (B, 4, 3, 3))
    u = rng.normal(0, 60, (B, 12 * N)); u[:, 2:12:3] = np.abs(u[:, 2:12:3]) + 30
    t_now = np.concatenate([1e-3 * rng.integers(0, 100000, B // 2), rng.uniform(0, 50, B - B // 2)])
    gait = ct.Gait(3.0, 0.6)
    tau, mask = ct.stance_torque(dev(J), dev(u), dev(t_now), gait, N, tau_max=45.0)
    tau, mask = tau.cpu().numpy(), mask.cpu().numpy()
    ref_mask = np.stack([traj_ref.current_mask(t, 3.0, 0.6) for t in t_now])
    assert np.array_equal(mask, ref_mask)
    f = u[:, :12].reshape(B, 4, 3)
    ref = np.clip(np.einsum("blij,bli->blj", J, -f), -45.0, 45.0) * ref_mask[:, :, None]     # J^T (-f), clipped
    assert np.abs(tau.reshape(B, 4, 3) - ref).max() < 1e-12
    assert (np.abs(tau) == 45.0).any() and (tau.reshape(B, 4, 3)[ref_mask == 0] == 0).all()


def test_leg_jacobian_feeds_the_stance_torque(mod):
    """cmpc_leg_jacobian (the analytic counterpart of compute_3x3_foot_Jacobian_world,
    go2_robot_data.py:286-300) against the oracle's axis-cross-lever Jacobian, then tau = clip(J^T (-f)) through
    cmpc_stance_torque (leg_controller.py:100-101)."""
    from convex_mpc_b200 import com_trajectory as ct
    from oracle import leg_kin
    rng = np.random.default_rng(21)
    B, N = 4096, 16
    q = np.stack([rng.uniform(-0.8, 0.8, (B, 4)), rng.uniform(-0.5, 2.0, (B, 4)), rng.uniform(-2.6, -0.9, (B, 4))], axis=2).reshape(B, 12)
    R = records._rot_zyx(rng.normal(0, 0.2, B), rng.normal(0, 0.2, B), rng.uniform(-np.pi, np.pi, B))
    R_wb = np.ascontiguousarray(np.swapaxes(R, 1, 2))
    J, pb = ct.leg_jacobian(dev(q), dev(R_wb), with_foot_pos=True)
    J, pb = J.cpu().numpy(), pb.cpu().numpy()
    for b in range(0, B, 37):
        for leg in range(4):
            q3 = q[b, 3 * leg:3 * leg + 3]
            assert np.abs(J[b, leg] - leg_kin.jacobian_world(q3, R_wb[b], leg_kin.SIDE[leg])).max() < 1e-14
            assert np.abs(pb[b, leg] - leg_kin.foot_pos_body(q3, leg_kin.SIDE[leg])).max() < 1e-14
    u = rng.normal(0, 60, (B, 12 * N)); u[:, 2:12:3] = np.abs(u[:, 2:12:3]) + 30
    t_now = rng.uniform(0, 50, B)
    tau, mask = ct.stance_torque(dev(J), dev(u), dev(t_now), ct.Gait(3.0, 0.6), N, tau_max=45.0)
    ref = np.clip(np.einsum("blij,bli->blj", J, -u[:, :12].reshape(B, 4, 3)), -45.0, 45.0) * mask.cpu().numpy()[:, :, None]
    assert np.abs(tau.cpu().numpy().reshape(B, 4, 3) - ref).max() < 1e-12


def test_drop_in_single_robot_api(mod):
    """The reference call pattern (test_MPC.py:153-192) with un-batched NumPy fields and Ad/Bd/gd."""
    rec = records.random_records(1, seed=9, stress=1.0)
    ct, Ad, Bd, gd = oracle_inputs(rec, 0)

    class Traj:       # what ComTraj exposes, reference shapes
        N = rec.N
        initial_x_vec = rec.x0[0].reshape(12, 1)
        contact_table = ct
        def compute_x_ref_vec(self):
            return rec.x_ref[0]
    traj = Traj()
    traj.Ad, traj.Bd, traj.gd = Ad, Bd, gd
    mpc = mod.CentroidalMPC(None, traj, verbose=False)
    sol = mpc.solve_QP(None, traj, False)
    w_opt = sol["x"].full().flatten()
    N = traj.N
    assert w_opt.shape == (24 * N,)
    X_opt = w_opt[:12 * N].reshape((12, N), order="F")
    U_opt = w_opt[12 * N:].reshape((12, N), order="F")
    o = oracle_solution(rec, 0)
    assert force_error(U_opt.reshape(-1, order="F"), o["sol"]["U"])[1] < 1e-3
    assert np.abs(X_opt.reshape(-1, order="F") - condensed_qp.rollout(o["cq"], o["sol"]["U"])).max() < 1e-8
    assert mpc.solve_time > 0 and mpc.update_time > 0
    assert sol["lam_x"].full().shape == (24 * N, 1) and sol["lam_a"].full().shape == (28 * N, 1)
    assert tuple(sol["u"].shape) == (12, N)
    # second call warm-starts from the first (centroidal_mpc.py:92-95) and returns the same optimum
    sol2 = mpc.solve_QP(None, traj, False)
    assert np.abs(sol2["x"].full().flatten() - w_opt).max() < 1e-7


def test_reference_style_traj_takes_fast_path(mod):
    """A traj shaped like the reference's ComTraj: un-batched NumPy fields, Ad/Bd/gd AND the raw members
    (m, I_com_world, r_*_foot_world) but no stored time step.  'auto' must pick the fast kernel (dt read from
    A_d[0,6]) and agree with the Ad/Bd route."""
    rec = records.random_records(1, seed=19, stress=0.6)
    ct, Ad, Bd, gd = oracle_inputs(rec, 0)

    class Traj:
        N = rec.N
        initial_x_vec = rec.x0[0].reshape(12, 1)
        contact_table = ct
        m = float(rec.mass[0])
        I_com_world = rec.I_world[0]
        r_fl_foot_world, r_fr_foot_world, r_rl_foot_world, r_rr_foot_world = (rec.r_foot[0, i] for i in range(4))
        def compute_x_ref_vec(self):
            return rec.x_ref[0]
    traj = Traj()
    traj.Ad, traj.Bd, traj.gd = Ad, Bd, gd
    fast = mod.CentroidalMPC(None, traj, verbose=False)
    slow = mod.CentroidalMPC(None, traj, verbose=False, dynamics="traj")
    a = fast.solve_QP(None, traj, False)
    b = slow.solve_QP(None, traj, False)
    wa, wb = a["x"].full().flatten(), b["x"].full().flatten()
    assert np.abs(wa - wb).max() < 1e-6
    o = oracle_solution(rec, 0)
    assert force_error(wa[12 * rec.N:], o["sol"]["U"])[1] < 1e-3
    assert int(a["status"]) == 1 and int(b["status"]) == 1


def test_AdBd_path_equals_device_dynamics_path(mod):
    rec = records.random_records(32, seed=77, stress=0.5)
    mpc, traj = make_mpc(mod, rec)
    a = mpc.solve_QP(None, traj)["u"].cpu().numpy()
    AB = [oracle_inputs(rec, b) for b in range(rec.B)]
    traj2 = mod.BatchedComTraj(rec.N, dev(rec.x0), dev(rec.x_ref), rec.dt, Ad=dev(np.stack([x[1] for x in AB])),
                               Bd=dev(np.stack([x[2] for x in AB])), gd=dev(np.stack([x[3].reshape(12) for x in AB])),
                               contact_table=dev(np.stack([x[0] for x in AB])))
    mpc2 = mod.CentroidalMPC(None, traj2, verbose=False)
    b = mpc2.solve_QP(None, traj2)["u"].cpu().numpy()
    assert np.abs(a - b).max() < 1e-6


def test_fast_kernel_equals_generic_kernel(mod):
    """Raw inputs through the closed-form fast kernel and through the generic (recursion) kernel."""
    rec = records.random_records(256, seed=123, stress=0.4)
    mpc_f, traj = make_mpc(mod, rec)
    mpc_g, _ = make_mpc(mod, rec, generic_kernel=True)
    a = mpc_f.solve_QP(None, traj)
    b = mpc_g.solve_QP(None, traj)
    assert (a["status"].cpu().numpy() == 1).all() and (b["status"].cpu().numpy() == 1).all()
    assert float((a["u"] - b["u"]).abs().max()) < 1e-6
    assert float((a["x"].tensor - b["x"].tensor).abs().max()) < 1e-6
    assert float((a["lam_a"].tensor - b["lam_a"].tensor).abs().max()) < 1e-6
    assert float((a["lam_x"].tensor - b["lam_x"].tensor).abs().max()) < 1e-6
    sa, sb = a["stats"].cpu().numpy(), b["stats"].cpu().numpy()
    assert np.abs(sa[:, 2] - sb[:, 2]).max() < 1e-8 * max(1.0, np.abs(sa[:, 2]).max())
    assert np.array_equal(sa[:, 3], sb[:, 3])


@pytest.mark.parametrize("N,B", [(32, 64), (48, 32)])
def test_longer_horizons(mod, N, B):
    """BASELINE configs[3]: N = 32 / 48 (384 / 576 condensed variables); the factor lives in L2-resident
    global scratch instead of shared memory."""
    rec = records.random_records(B, N=N, seed=16384 + N, stress=0.2)
    mpc, traj = make_mpc(mod, rec, max_stance=4 * (int(0.6 * N) + 1))
    sol = mpc.solve_QP(None, traj)
    assert (sol["status"].cpu().numpy() == 1).all()
    u = sol["u"].cpu().numpy()
    for b in range(0, B, max(1, B // 4)):
        o = oracle_solution(rec, b)
        assert force_error(u[b].reshape(-1, order="F"), o["sol"]["U"])[1] < 1.0


def test_closed_loop_replay_1024_robots(mod):
    """BASELINE configs[1]: 1024 robots, mixed forward / lateral / yaw commands, replayed in closed loop with
    the warm start the reference uses (previous solution, unshifted, centroidal_mpc.py:92-95).  Every cycle:
    all QPs solved and self-certified, sampled robots match the oracle's exact optimum, and the warm-started
    solve equals a cold solve of the same cycle."""
    rec = records.random_records(1024, seed=1024)
    traj = mod.BatchedComTraj.from_records(rec, device="cuda:0")
    mpc = mod.CentroidalMPC(None, traj, verbose=False, max_stance=40)
    cold = mod.CentroidalMPC(None, traj, verbose=False, max_stance=40)
    z0 = rec.x0[:, 2].copy()
    for cycle in range(12):
        traj = mod.BatchedComTraj.from_records(rec, device="cuda:0")
        sol = mpc.solve_QP(None, traj)                      # warm-started from cycle - 1
        st = sol["stats"].cpu().numpy()
        assert (sol["status"].cpu().numpy() == 1).all(), cycle
        assert st[:, 0].max() < 1e-7 and st[:, 1].max() < 1e-7
        u = sol["x"].full()[:, 12 * rec.N:]
        cold.reset()
        uc = cold.solve_QP(None, traj)["x"].full()[:, 12 * rec.N:]
        assert np.abs(u - uc).max() < 1e-6
        for b in (3, 511, 1000):
            o = oracle_solution(rec, b)
            assert force_error(u[b], o["sol"]["U"])[1] < 1e-3, (cycle, b)
        rec = records.next_cycle(rec, u[:, :12])
    # the fleet is still standing and moving: height kept, forward speed near the command on average
    assert np.abs(rec.x0[:, 2] - z0).max() < 0.05
    assert np.abs(rec.x0[:, 3:5]).max() < 0.5


def test_admm_mode_and_polish(mod):
    rec = records.random_records(32, seed=31, stress=0.3)
    mpc, traj = make_mpc(mod, rec, mode="admm", eps_abs=1e-5, eps_rel=1e-5, max_iter=4000)
    sol = mpc.solve_QP(None, traj)
    st = sol["stats"].cpu().numpy()
    assert (sol["status"].cpu().numpy() == 1).all() and (sol["iters"].cpu().numpy() > 0).all()
    assert (st[:, 7] == 2).all()
    u_admm = sol["u"].cpu().numpy()
    mpc2, _ = make_mpc(mod, rec, mode="admm", polish=True)
    u_pol = mpc2.solve_QP(None, traj)["u"].cpu().numpy()
    for b in range(rec.B):
        U = oracle_solution(rec, b)["sol"]["U"]
        assert np.abs(u_admm[b].reshape(-1, order="F") - U).max() < 2.0      # OSQP-like: loose in flat directions
        assert force_error(u_pol[b].reshape(-1, order="F"), U)[1] < 1.0


def test_edge_cases(mod):
    rec = records.random_records(8, seed=41, stress=0.5)
    N = rec.N
    for val, nfree in ((0, 0), (1, 12 * N)):
        traj = mod.BatchedComTraj.from_records(rec, device="cuda:0")
        traj.contact_table = torch.full((rec.B, 4, N), val, dtype=torch.int32, device="cuda")
        mpc = mod.CentroidalMPC(None, traj, verbose=False)
        sol = mpc.solve_QP(None, traj)
        st = sol["stats"].cpu().numpy()
        assert (sol["status"].cpu().numpy() == 1).all() and (st[:, 3] == nfree).all()
        u = sol["u"].cpu().numpy()
        for b in range(3):
            o = oracle_solution(rec, b, contact=np.full((4, N), val))
            assert force_error(u[b].reshape(-1, order="F"), o["sol"]["U"])[1] < 1e-3
    # B = 1 batched, and a stance bound that is too small is reported per robot
    one = rec.slice(0, 1)
    mpc, traj = make_mpc(mod, one)
    assert tuple(mpc.solve_QP(None, traj)["u"].shape) == (1, 12, N)
    mpc, traj = make_mpc(mod, rec, max_stance=8)
    assert (mpc.solve_QP(None, traj)["status"].cpu().numpy() == -20).all()


def test_host_entry_matches_device_entry(mod):
    rec = records.random_records(3000, seed=88, stress=0.2)
    mpc, traj = make_mpc(mod, rec)
    a = mpc.solve_QP(None, traj)["x"].full()[:, 12 * rec.N:]
    u, st, it = mpc.solve_host(rec.x0, rec.x_ref, rec.r_foot, rec.I_world, rec.mass, rec.t0, rec.dt, rec.gait_hz, rec.duty)
    assert (st.numpy() == 1).all()
    assert np.abs(u.numpy() - a).max() < 1e-9


def test_full_size_properties(mod):
    """BASELINE config #3 at full size (65 536 robots): every QP certified by its own KKT residuals
    (computed in-kernel from the roll-out and co-states, independent of the factorisation)."""
    rec = records.random_records(65536, seed=65536)
    mpc, traj = make_mpc(mod, rec, max_stance=40)
    sol = mpc.solve_QP(None, traj)
    st = sol["stats"].cpu().numpy()
    status = sol["status"].cpu().numpy()
    assert (status == 1).all()
    assert st[:, 0].max() < 1e-8 and st[:, 1].max() < 1e-8
    u = sol["u"]
    # swing legs carry exactly zero force; stance legs respect fz >= 10 and the pyramid
    ct = torch.from_numpy(records.host_contact_table(rec.t0, rec.dt, rec.N, rec.gait_hz, rec.duty)).cuda()
    f = u.reshape(rec.B, 4, 3, rec.N)
    swing = (ct == 0)
    assert float(f.abs().amax(dim=2)[swing].max()) == 0.0
    fz = f[:, :, 2, :][~swing]
    assert float(fz.min()) >= 10.0 - 1e-8
    assert float((f[:, :, 0, :].abs()[~swing] - 0.8 * fz).max()) <= 1e-8
    assert float((f[:, :, 1, :].abs()[~swing] - 0.8 * fz).max()) <= 1e-8
    # total vertical force at the first step ~ weight (sanity of the whole pipeline)
    fz0 = f[:, :, 2, 0].sum(dim=1)
    assert 0.3 * 15.02 * 9.81 < float(fz0.median()) < 3.0 * 15.02 * 9.81
    # idempotence: a warm-started second solve returns the same forces
    u1 = u.clone()
    u2 = mpc.solve_QP(None, traj)["u"]
    assert float((u1 - u2).abs().max()) < 1e-6
    # sampled L2 check against the oracle
    un = u1.cpu().numpy()
    for b in range(0, rec.B, 2048):
        o = oracle_solution(rec, b)
        assert force_error(un[b].reshape(-1, order="F"), o["sol"]["U"])[1] < 1e-3


def test_automatic_stance_bound(mod):
    """Without ``max_stance`` the host class derives the periodic-gait bound itself (40 foot-steps at N = 16, 3 Hz / 0.6)
    and falls back to 4N as soon as a caller-supplied contact table arrives; the solutions do not depend on the bound."""
    rec = records.random_records(512, seed=77, stress=0.3)
    auto, traj = make_mpc(mod, rec)
    assert auto._auto_stance == 40
    full, _ = make_mpc(mod, rec, max_stance=4 * rec.N)
    assert full._auto_stance is None
    a, b = auto.solve_QP(None, traj), full.solve_QP(None, traj)
    assert (a["status"].cpu().numpy() == 1).all() and (b["status"].cpu().numpy() == 1).all()
    assert float((a["u"] - b["u"]).abs().max()) < 1e-7
    # a table with every leg in stance over the whole horizon (64 foot-steps) must not be refused
    tr2 = mod.BatchedComTraj.from_records(rec, device="cuda:0", with_contact_table=True)
    tr2.contact_table = torch.ones_like(tr2.contact_table)
    c = auto.solve_QP(None, tr2)
    assert auto._auto_stance is None
    assert (c["status"].cpu().numpy() == 1).all()
