"""Host-side workload utilities (no GPU): the recorded-states replay format of BASELINE configs[1],
the command schedule helper and the per-cycle inputs of cmpc_cycle_host."""
import numpy as np

from convex_mpc_b200 import records


def test_save_and_load_cycles_round_trip(tmp_path):
    rec = records.random_records(8, seed=3, stress=0.2)
    rec = records.retarget(rec, np.linspace(-0.8, 0.8, 8), 0.1, np.linspace(-4, 4, 8))
    nxt = records.next_cycle(rec, np.tile([0.0, 0.0, 36.8], 4)[None].repeat(8, axis=0))
    u = [np.arange(8 * 12 * rec.N, dtype=np.float64).reshape(8, -1), np.ones((8, 12 * rec.N))]
    path = str(tmp_path / "cycles.npz")
    records.save_cycles(path, [rec, nxt], u)
    back, ub = records.load_cycles(path)
    assert len(back) == 2 and back[0].N == rec.N and back[1].dt == rec.dt and back[0].gait_hz == rec.gait_hz
    for a, b in zip(back, (rec, nxt)):
        for k in ("x0", "x_ref", "r_foot", "I_world", "mass", "t0"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(ub[0], u[0]) and np.array_equal(ub[1], u[1])
    # without forces
    records.save_cycles(path, [rec])
    assert records.load_cycles(path)[1] is None


def test_retarget_rebuilds_the_reference_window():
    """The reference restarts the window from the current position with the commanded body-frame velocity rotated by
    the current yaw (com_trajectory.py:84-103)."""
    rec = records.random_records(5, seed=11)
    out = records.retarget(rec, 0.5, -0.2, 1.5)
    N, dt = rec.N, rec.dt
    yaw = rec.x0[:, 5]
    vx = np.cos(yaw) * 0.5 + np.sin(yaw) * 0.2
    assert np.allclose(out.x_ref[:, 6, :], vx[:, None]) and np.allclose(out.x_ref[:, 11, :], 1.5)
    assert np.allclose(out.x_ref[:, 5, -1], yaw + 1.5 * N * dt)
    assert np.allclose(out.x_ref[:, 0, 0], rec.x0[:, 0] + vx * dt) and np.allclose(out.x_ref[:, 2, :], 0.27)
    assert out.r_foot is rec.r_foot and np.array_equal(out.x0, rec.x0)


def test_cycle_inputs_are_consistent():
    g = records.random_cycle_inputs(32, seed=5)
    assert g["x0"].shape == (32, 12) and g["R_wb"].shape == (32, 3, 3) and g["lever"].shape == (32, 4, 3) and g["cmd"].shape == (32, 4)
    # R_world_to_body is a rotation and I_world = R I_body R'
    R = np.swapaxes(g["R_wb"], 1, 2)
    assert np.allclose(np.einsum("bij,bkj->bik", R, R), np.eye(3)[None], atol=1e-12)
    assert np.allclose(g["I_world"], np.einsum("bij,j,bkj->bik", R, records.GO2_I_BODY, R))
    assert np.all(g["lever"][:, :, 2] < 0) and np.allclose(g["pos_des"], g["x0"][:, 0:3])
