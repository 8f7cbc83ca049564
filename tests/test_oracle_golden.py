"""Pin the oracle restatement against outputs of the reference's own code (tests/golden/).

Rows a1-a4 of SURVEY.md section 8: contact tables bit-exact, dynamics to ~1e-16.
"""
import numpy as np

from oracle import dynamics_ref, gait_ref


def test_contact_tables_bit_exact(golden):
    for ci in range(int(golden["ct_count"])):
        hz, duty, N, dt = golden[f"ct{ci}_cfg"]
        N = int(N)
        t0 = golden[f"ct{ci}_t0"]
        ref = golden[f"ct{ci}_table"]
        hz_arg = int(hz) if float(hz).is_integer() else float(hz)
        for i, t in enumerate(t0):
            got = gait_ref.contact_table(float(t), float(dt), N, hz_arg, float(duty))
            assert got.dtype == np.int32 and got.shape == (4, N)
            assert np.array_equal(got, ref[i]), (ci, i, t)


def test_mask_pack_roundtrip(golden):
    tab = golden["ct0_table"]
    words = gait_ref.pack_mask(tab)
    assert words.shape == (tab.shape[0], 1) and words.dtype == np.uint64
    assert np.array_equal(gait_ref.unpack_mask(words, 16), tab)
    tab48 = golden["ct2_table"]
    w48 = gait_ref.pack_mask(tab48)
    assert w48.shape[-1] == 3
    assert np.array_equal(gait_ref.unpack_mask(w48, 48), tab48)


def test_dynamics_match_reference(golden):
    for ci in range(int(golden["dyn_count"])):
        N, dt, m = golden[f"dyn{ci}_in_scalar"]
        N = int(N)
        xref = golden[f"dyn{ci}_xref"]
        yaw = dynamics_ref.yaw_average(xref)
        Ac, Bc, gc = dynamics_ref.continuous_dynamics(m, golden[f"dyn{ci}_I"], yaw, golden[f"dyn{ci}_rfoot"])
        assert np.array_equal(Ac, golden[f"dyn{ci}_Ac"])
        assert np.abs(Bc - golden[f"dyn{ci}_Bc"]).max() <= 1e-15
        for fn in (dynamics_ref.discrete_dynamics_closed, dynamics_ref.discrete_dynamics_literal):
            Ad, Bd, gd = fn(Ac, Bc, gc, dt)
            assert np.abs(Ad - golden[f"dyn{ci}_Ad"]).max() <= 1e-16
            assert np.abs(Bd - golden[f"dyn{ci}_Bd"]).max() <= 1e-15
            assert np.abs(gd - golden[f"dyn{ci}_gd"]).max() <= 1e-16
            assert gd.shape == (12, 1)


def test_traj_generator_restatement_matches_reference():
    """oracle/traj_ref.py against the reference's own generate_traj (tests/golden/make_golden_traj.py)."""
    from helpers import golden_traj_batches
    from oracle import traj_ref
    n = 0
    for g in golden_traj_batches():
        for b in range(g["x0"].shape[0]):
            pd, xr, rf = traj_ref.generate_traj(g["x0"][b], g["R_wb"][b], g["levers"][b], g["cmd"][b], float(g["t_now"][b]),
                                                g["dt"], g["N"], g["hz"], g["duty"], g["hip"], g["pos_des_in"][b])
            assert np.array_equal(pd, g["pos_des_out"][b])
            assert np.array_equal(xr, g["x_ref"][b])
            assert np.array_equal(rf == 0.0, g["r_foot"][b] == 0.0)            # swing / stance pattern: exact
            assert np.abs(rf - g["r_foot"][b]).max() <= 1e-15
            n += 1
    assert n == 72
