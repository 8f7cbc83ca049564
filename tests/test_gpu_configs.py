"""Gated GPU tests (-m gpu) of the BASELINE.json configurations at their full sizes, through the drop-in class:

  configs[3]  horizon sweep N = 16 / 32 / 48 at batch 16 384, tolerance vs the exact optimum
  configs[4]  262 144 robots (one GPU here; the sharded run is bench.py --gpus N)
  north star  "OSQP-equivalent ADMM": OSQP's termination inequalities at eps 1e-5 on >= 1 024 QPs, polish within tolerance
  host entry  cmpc_solve_host (chunked, two streams) against the device entry on forces, nominal / disturbed / N = 32

Size-independent properties carry the full sizes (every QP self-certified: KKT residuals from the roll-out and the
co-states; swing forces exactly zero; fz >= fz_min; pyramid), sampled robots are checked against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from convex_mpc_b200 import records  # noqa: E402
from helpers import force_error, oracle_solution  # noqa: E402
from oracle import sparse_qp  # noqa: E402

TOL_ABS, TOL_REL = 1e-2, 1e-3      # BASELINE.json north_star


@pytest.fixture(scope="module")
def mod():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from convex_mpc_b200 import centroidal_mpc
    return centroidal_mpc


def stance_bound(rec):
    return 4 * (int(np.floor(rec.duty * rec.N)) + 1)


def check_batch(mod, rec, n_oracle, **kw):
    traj = mod.BatchedComTraj.from_records(rec, device="cuda:0")
    mpc = mod.CentroidalMPC(None, traj, verbose=False, max_stance=stance_bound(rec), **kw)
    sol = mpc.solve_QP(None, traj)
    torch.cuda.synchronize()
    status = sol["status"].cpu().numpy()
    st = sol["stats"].cpu().numpy()
    assert (status == 1).all(), np.unique(status, return_counts=True)
    fb = np.isin(st[:, 7], (2, 3))                    # the condensed kernel's ADMM fallback stops at its 1e-6 tolerance
    assert st[~fb, 0].max() < 1e-8 and st[~fb, 1].max() < 1e-7               # primal / dual residual of every exact QP
    assert fb.mean() < 1e-3 and (not fb.any() or (st[fb, 0].max() < 1e-3 and st[fb, 1].max() < 1e-3))
    u = sol["u"]                                                             # (B,12,N)
    N, B = rec.N, rec.B
    ct = torch.from_numpy(records.host_contact_table(rec.t0, rec.dt, N, rec.gait_hz, rec.duty)).cuda()
    f = u.reshape(B, 4, 3, N)
    swing = (ct == 0)
    assert float(f.abs().amax(dim=2)[swing].max()) == 0.0                   # centroidal_mpc.py:150-161
    # feasibility per robot: exact paths to 1e-8, the few ADMM-fallback robots to their 1e-3 N stop (tolerance 1e-2 N)
    tol = torch.from_numpy(np.where(fb, 1e-3, 1e-8)).cuda()[:, None, None]
    big = torch.full_like(f[:, :, 2, :], 1e9)
    fz = f[:, :, 2, :]
    assert bool(((torch.where(swing, big, fz) - 10.0) >= -tol).all())        # :163-170
    for c in (0, 1):                                                         # :324-359
        assert bool((torch.where(swing, -big, f[:, :, c, :].abs() - 0.8 * fz) <= tol).all())
    un = u.cpu().numpy()
    worst = 0.0
    for b in np.linspace(0, B - 1, n_oracle).astype(int):
        o = oracle_solution(rec, b)
        err, rel = force_error(un[b].reshape(-1, order="F"), o["sol"]["U"])
        assert rel < 1.0, (b, err)
        worst = max(worst, err)
    return st, worst


@pytest.mark.parametrize("N", [16, 32, 48])
def test_config3_horizon_sweep_batch_16384(mod, N):
    rec = records.random_records(16384, N=N, seed=16384 + N)
    st, worst = check_batch(mod, rec, n_oracle=16 if N == 16 else 8)
    assert worst < 1e-6
    paths = np.bincount(st[:, 7].astype(int), minlength=6)
    assert paths[4] + paths[5] > 0.95 * rec.B               # O(N) route: Riccati sweeps, not the dense condensed factor


def test_config4_262144_robots(mod):
    rec = records.random_records(262144, seed=262144)
    st, worst = check_batch(mod, rec, n_oracle=64)
    assert worst < 1e-6


def test_disturbed_batch_65536(mod):
    """30 % of the robots shoved hard: every QP still solved exactly (no max_iter, no inaccurate)."""
    rec = records.random_records(65536, seed=65536, stress=0.3)
    st, worst = check_batch(mod, rec, n_oracle=16)
    exact = ~np.isin(st[:, 7], (2,))
    assert exact.mean() > 0.999


def test_admm_mode_meets_osqp_termination_on_1024_qps(mod):
    """mode='admm' (the OSQP-equivalent solver of the north star) at eps 1e-5: the lifted point (w, lam_x, lam_a) passes
    OSQP's termination inequalities on the reference's sparse QP; with polish the forces are within the tolerance."""
    rec = records.random_records(1024, seed=1024, stress=0.2)
    traj = mod.BatchedComTraj.from_records(rec, device="cuda:0")
    eps = 1e-5
    mpc = mod.CentroidalMPC(None, traj, verbose=False, mode="admm", eps_abs=eps, eps_rel=eps, max_iter=4000)
    sol = mpc.solve_QP(None, traj)
    assert (sol["status"].cpu().numpy() == 1).all() and (sol["stats"].cpu().numpy()[:, 7] == 2).all()
    w = sol["x"].full()
    lam_x, lam_a = sol["lam_x"].full(), sol["lam_a"].full()
    pol = mod.CentroidalMPC(None, traj, verbose=False, mode="admm", eps_abs=eps, eps_rel=eps, max_iter=4000, polish=True)
    up = pol.solve_QP(None, traj)["x"].full()[:, 12 * rec.N:]
    from helpers import oracle_inputs
    N = rec.N
    for b in range(rec.B):
        ct, Ad, Bd, gd = oracle_inputs(rec, b)
        sq = sparse_qp.build(Ad, Bd, gd, rec.x0[b], rec.x_ref[b], ct)
        P, q, A, l, uu = sparse_qp.as_osqp_form(sq)
        y = np.concatenate([lam_x[b], lam_a[b]])
        Ax = A @ w[b]
        z = np.clip(Ax, l, uu)
        r_p = np.abs(Ax - z).max()
        r_d = np.abs(P @ w[b] + q + A.T @ y).max()
        assert r_p <= eps + eps * max(np.abs(Ax).max(), np.abs(z).max()), (b, r_p)
        assert r_d <= eps + eps * max(np.abs(P @ w[b]).max(), np.abs(A.T @ y).max(), np.abs(q).max()), (b, r_d)
    for b in range(0, rec.B, 16):
        U = oracle_solution(rec, b)["sol"]["U"]
        d = np.abs(up[b] - U)
        assert (d <= TOL_ABS + TOL_REL * np.abs(U)).all(), (b, d.max())


@pytest.mark.parametrize("N,stress", [(16, 0.0), (16, 0.3), (32, 0.0)])
def test_host_entry_equals_device_entry_multichunk(mod, N, stress):
    """cmpc_solve_host at sizes that take several chunks on both streams, repeated to flush scheduling-dependent races
    (every in-flight solve owns its workspace slot)."""
    B = 65536 if N == 16 else 32768
    rec = records.random_records(B, N=N, seed=7 + N, stress=stress)
    traj = mod.BatchedComTraj.from_records(rec, device="cuda:0")
    mpc = mod.CentroidalMPC(None, traj, verbose=False, max_stance=stance_bound(rec), max_batch=B)
    a = mpc.solve_QP(None, traj)
    ua = a["x"].full()[:, 12 * N:]
    exact = ~np.isin(a["stats"].cpu().numpy()[:, 7], (2, 3))
    for rep in range(3):
        mpc._warm_host = 0
        u, st, it = mpc.solve_host(rec.x0, rec.x_ref, rec.r_foot, rec.I_world, rec.mass, rec.t0, rec.dt, rec.gait_hz, rec.duty)
        assert (st.numpy() == 1).all()
        assert np.abs(u.numpy() - ua)[exact].max() < 1e-9, rep


@pytest.mark.parametrize("first_only", [False, True])
def test_cycle_host_matches_device_pipeline(mod, first_only):
    """cmpc_cycle_host (408 bytes in per robot, trajectory generated on the device, optionally only U_opt[:, 0] out --
    test_MPC.py:173-196) against generate_traj -> solve_QP on device tensors: same forces, same carried position target."""
    from convex_mpc_b200 import com_trajectory as ct
    B, N, HZ, DUTY = 20000, 16, 3.0, 0.6
    g = records.random_cycle_inputs(B, 2027)
    gait = ct.Gait(HZ, DUTY)
    dt = gait.gait_period / N
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    state = ct.RobotState(d(g["x0"]), d(g["R_wb"]), d(g["lever"]), d(g["mass"]), d(g["I_world"]))
    traj = ct.ComTraj(state, hip_offset=g["hip"], device="cuda:0")
    traj.pos_des_world.copy_(d(g["pos_des"]))
    traj.generate_traj(state, gait, d(g["t0"]), *[d(g["cmd"][:, i]) for i in range(4)], dt)
    mpc = mod.CentroidalMPC(None, traj, verbose=False, max_stance=40)
    sol = mpc.solve_QP(None, traj)
    u_dev = mpc._u.cpu().numpy()
    host = mod.CentroidalMPC(None, traj, verbose=False, max_stance=40, max_batch=B)
    pos_des = g["pos_des"].copy()
    u, st, it = host.cycle_host(g["x0"], g["R_wb"], g["lever"], g["cmd"], g["t0"], pos_des, g["I_world"], g["mass"], dt, g["hip"],
                                gait_hz=HZ, duty=DUTY, first_step_only=first_only)
    u, st = u.numpy(), st.numpy()
    assert (st == 1).all() and (sol["status"].cpu().numpy() == 1).all()
    assert np.array_equal(pos_des, traj.pos_des_world.cpu().numpy())
    ref = u_dev[:, :12] if first_only else u_dev
    assert u.shape == ref.shape and np.abs(u - ref).max() < 1e-9
    # a second, warm-started cycle through the host entry keeps working from the state left on the device
    u2, st2, _ = host.cycle_host(g["x0"], g["R_wb"], g["lever"], g["cmd"], g["t0"], pos_des, g["I_world"], g["mass"], dt, g["hip"],
                                 gait_hz=HZ, duty=DUTY, first_step_only=first_only)
    assert (st2.numpy() == 1).all() and np.abs(u2.numpy() - ref).max() < 1e-6


def test_config1_record_and_replay(mod, tmp_path):
    """BASELINE configs[1] in miniature: 64 robots x 12 closed-loop cycles recorded to the .npz replay format
    (records.save_cycles), then replayed in order, warm-started: the same forces come back."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import replay as rp
    path = str(tmp_path / "replay_64x12.npz")
    cyc, us = rp.record(64, 12, path)
    back, u_back = records.load_cycles(path)
    assert len(back) == 12 and back[0].B == 64 and np.array_equal(back[5].x_ref, cyc[5].x_ref) and np.array_equal(u_back[7], us[7])
    res = rp.replay(path, str(tmp_path / "replay.json"))
    assert res["max_abs_force_difference_to_recording_N"] < 1e-6
