"""GPU parity tests (-m gpu) of the wrench-space projected-Riccati active-set kernel (csrc/cmpc_wrench.cuh, pre-pass 4,
the default route of ``cmpc_solve`` for batches of 2 048 robots and more): same optimum as the condensed active-set
kernel (pre-pass 0) and as the oracle's exact solution, outputs in the reference's layouts (centroidal_mpc.py:69-120,
consumer contract test_MPC.py:189-196)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from convex_mpc_b200 import records  # noqa: E402
from helpers import Emul, force_error, oracle_solution  # noqa: E402
from oracle import condensed_qp, wrench_riccati  # noqa: E402


@pytest.fixture(scope="module")
def mod():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from convex_mpc_b200 import centroidal_mpc
    return centroidal_mpc


def solve(mod, rec, **kw):
    traj = mod.BatchedComTraj.from_records(rec, device="cuda:0")
    mpc = mod.CentroidalMPC(None, traj, verbose=False, **kw)
    mpc.solve_QP(None, traj)
    torch.cuda.synchronize()
    out = dict(u=mpc._u.cpu().numpy(), y=mpc._y.cpu().numpy(), X=mpc._X.cpu().numpy(), nu=mpc._nu.cpu().numpy(),
               status=mpc._status.cpu().numpy(), stats=mpc._stats.cpu().numpy())
    return mpc, traj, out


@pytest.mark.parametrize("stress,N", [(0.0, 16), (0.3, 16), (1.0, 16), (0.0, 32), (0.3, 48)])
def test_same_optimum_as_condensed_kernel_and_oracle(mod, stress, N):
    B = 4096 if N == 16 else 2048
    rec = records.random_records(B, N=N, seed=100 + N + int(10 * stress), stress=stress)
    _, _, a = solve(mod, rec, prepass=4)
    _, _, b = solve(mod, rec, prepass=0)
    assert (a["status"] == 1).all() and (b["status"] == 1).all()
    path = a["stats"][:, 7].astype(int)
    wrench = np.isin(path, (4, 5))
    # cycling working sets are moved off the cycle by single exchanges; only oscillating block updates of heavily
    # disturbed robots (40+ active rows) and spent budgets go on to the condensed kernel
    assert wrench.mean() > (0.999 if stress == 0 else (0.99 if stress < 1.0 else 0.95))
    if stress > 0:
        assert (path == 5).mean() > 0.1                            # constrained robots are finished here too
    exact_b = ~np.isin(b["stats"][:, 7], (2,))              # the condensed kernel's ADMM fallback stops at its 1e-6 tolerance
    # rows the condensed kernel left to its ADMM fallback are only as close as that fallback's 1e-6 stop leaves them
    # on this flat QP (a few 0.1 N); the wrench kernel's own certificate below and the oracle sample cover them
    assert np.abs(a["u"] - b["u"])[exact_b].max() < 1e-6 and np.abs(a["u"] - b["u"]).max() < 1.0
    assert a["stats"][wrench, 0].max() < 1e-9 and a["stats"][wrench, 1].max() < 1e-8  # own KKT certificate
    assert np.abs(a["stats"][:, 2] - b["stats"][:, 2]).max() < 1e-6 * np.abs(b["stats"][:, 2]).max()
    assert np.array_equal(a["stats"][:, 3], b["stats"][:, 3])
    assert np.array_equal(a["stats"][wrench & exact_b, 4], b["stats"][wrench & exact_b, 4])     # same number of active rows
    assert np.abs(a["X"] - b["X"])[exact_b].max() < 1e-9 and np.abs(a["nu"] - b["nu"])[exact_b].max() < 1e-6
    assert np.abs(a["y"] - b["y"])[exact_b].max() < 1e-6
    # oracle: exact optimum, independent certificate, lifted duals (a sample that contains constrained robots)
    pick = list(np.flatnonzero(path == 5)[:5]) + list(np.flatnonzero(path == 4)[:3])
    for i in pick:
        sol = oracle_solution(rec, i)
        err, rel = force_error(a["u"][i], sol["sol"]["U"])
        assert err < 1e-7, (i, err)
        k = condensed_qp.kkt_residuals(sol["cq"], a["u"][i], a["y"][i])
        assert k["stat"] < 1e-9 and k["prim"] < 1e-9 and k["dual"] < 1e-9 and k["comp"] < 1e-7
        w, lam_x, lam_a = condensed_qp.lift(sol["cq"], a["u"][i], a["y"][i], Ad=sol["Ad"], Bd=sol["Bd"], x_ref=rec.x_ref[i])
        assert np.abs(a["X"][i] - w[:12 * N]).max() < 1e-10
        assert np.abs(a["nu"][i] - lam_a[:12 * N]).max() < 1e-8


def test_matches_host_emulation_and_numpy_twin(mod):
    """Same working-set path (number of sweeps) and forces as the one-quad host emulation of the same source and as
    oracle/wrench_riccati.py."""
    rec = records.random_records(2048, seed=77, stress=0.5)
    _, _, a = solve(mod, rec, prepass=4)
    sub = rec.slice(0, 48)
    em = Emul().wrench(sub)
    for i in range(sub.B):
        if not em["done"][i]:
            assert a["stats"][i, 7] not in (4, 5)
            continue
        assert a["stats"][i, 7] in (4, 5)
        assert a["stats"][i, 6] == em["sweeps"][i] - 1
        assert np.abs(a["u"][i] - em["u"][i]).max() < 1e-8
    for i in range(6):
        ct = np.asarray(oracle_solution(rec, i)["ct"])
        rb = wrench_riccati.Robot(rec.x0[i], rec.x_ref[i], rec.r_foot[i], rec.I_world[i], rec.mass[i], rec.dt, ct)
        tw = wrench_riccati.solve(rb)
        if tw["ok"]:
            assert np.abs(a["u"][i] - tw["U"]).max() < 1e-8


def test_warm_start_from_previous_duals(mod):
    """The working set of the previous solution is the first guess (centroidal_mpc.py:92-95): one sweep per robot."""
    rec = records.random_records(4096, seed=5, stress=0.3)
    mpc, traj, a = solve(mod, rec, prepass=4)
    mpc.solve_QP(None, traj)          # warm
    torch.cuda.synchronize()
    st = mpc._stats.cpu().numpy()
    u2 = mpc._u.cpu().numpy()
    ok = np.isin(a["stats"][:, 7], (4, 5))
    assert (mpc._status.cpu().numpy() == 1).all()
    assert np.abs(u2 - a["u"])[ok].max() < 1e-7
    assert (st[ok, 6] == 0).mean() > 0.99


def test_shifted_warm_start_same_optimum(mod):
    """warm_shift=True (the previous working set shifted by one horizon stage; the reference re-uses the
    unshifted previous solution, centroidal_mpc.py:108-110) is only a different first guess: the next cycle's optimum is
    the cold-start optimum, and the closed-loop successor of a batch needs no more sweeps than with the unshifted guess."""
    rec = records.random_records(4096, seed=31, stress=0.2)
    _, _, cold0 = solve(mod, rec, prepass=4)
    nxt = records.next_cycle(rec, cold0["u"][:, :12])
    _, _, cold = solve(mod, nxt, prepass=4)
    sweeps = {}
    for shift in (False, True):
        traj0 = mod.BatchedComTraj.from_records(rec, device="cuda:0")
        traj1 = mod.BatchedComTraj.from_records(nxt, device="cuda:0")
        mpc = mod.CentroidalMPC(None, traj0, verbose=False, prepass=4, warm_shift=shift)
        mpc.solve_QP(None, traj0)
        mpc.solve_QP(None, traj1)                 # warm
        torch.cuda.synchronize()
        st = mpc._stats.cpu().numpy()
        assert (mpc._status.cpu().numpy() == 1).all()
        ok = np.isin(st[:, 7], (4, 5)) & np.isin(cold["stats"][:, 7], (4, 5))
        assert ok.mean() > 0.99
        assert np.abs(mpc._u.cpu().numpy() - cold["u"])[ok].max() < 1e-6
        sweeps[shift] = float(1 + st[ok, 6].mean())
    assert sweeps[True] <= sweeps[False] + 0.05, sweeps


@pytest.mark.parametrize("N", [18, 20])
def test_horizons_off_the_four_stage_grid(mod, N):
    """The Riccati route fetches lever arms four stages at a time, so it serves horizons that are multiples of 4 (N = 20);
    any other horizon (N = 18) takes the round-1 route (lock-step pre-pass + condensed kernel).  Same exact optimum either way."""
    rec = records.random_records(2048, N=N, seed=900 + N, stress=0.2)
    _, _, a = solve(mod, rec, prepass=4)
    assert (a["status"] == 1).all()
    path = a["stats"][:, 7].astype(int)
    if N % 4 == 0:
        assert np.isin(path, (4, 5)).mean() > 0.99 and (path == 5).any()
    else:
        assert not (path == 5).any() and (path == 4).any()
    for i in list(np.flatnonzero(path != 4)[:3]) + list(np.flatnonzero(path == 4)[:2]):
        sol = oracle_solution(rec, i)
        err, rel = force_error(a["u"][i], sol["sol"]["U"])
        assert err < 1e-6, (i, err)
