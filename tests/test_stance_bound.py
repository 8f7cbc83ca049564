"""Host logic: the automatic bound on stance foot-steps (CentroidalMPC._stance_bound) against a brute-force count
over the contact tables the gait produces (gait.py:26-37 arithmetic in records.host_contact_table)."""
from types import SimpleNamespace

import numpy as np
import pytest

from convex_mpc_b200.centroidal_mpc import CentroidalMPC
from convex_mpc_b200.records import host_contact_table


@pytest.mark.parametrize("hz,duty,N", [(3.0, 0.6, 16), (3.0, 0.6, 32), (2.5, 0.5, 16), (2.0, 0.75, 12), (4.0, 0.3, 10),
                                        (3.0, 0.95, 16), (3.0, 0.05, 16)])
def test_bound_covers_every_phase(hz, duty, N):
    dt = (1.0 / hz) / N
    traj = SimpleNamespace(contact_table=None, time_now=0.0, gait_hz=hz, gait_duty=duty, dt=dt)
    bound = CentroidalMPC._stance_bound(SimpleNamespace(N=N), traj)
    assert bound is not None and bound <= 4 * N
    t0 = np.concatenate([np.linspace(0.0, 3.0 / hz, 20001), 1e-3 * np.arange(10000)])
    tab = host_contact_table(t0, dt, N, hz, duty)
    worst = int(tab.sum(axis=(1, 2)).max())
    assert worst <= bound
    assert bound - worst <= 4          # tight to within one sample per leg


def test_bound_declines_when_it_does_not_apply():
    me = SimpleNamespace(N=16)
    dt = (1.0 / 3.0) / 16
    ok = dict(contact_table=None, time_now=0.0, gait_hz=3.0, gait_duty=0.6, dt=dt)
    assert CentroidalMPC._stance_bound(me, SimpleNamespace(**ok)) == 40
    assert CentroidalMPC._stance_bound(me, SimpleNamespace(**{**ok, "contact_table": np.ones((4, 16))})) is None
    assert CentroidalMPC._stance_bound(me, SimpleNamespace(**{**ok, "time_now": None})) is None
    assert CentroidalMPC._stance_bound(me, SimpleNamespace(**{**ok, "dt": 0.03})) is None      # horizon != one period
    assert CentroidalMPC._stance_bound(me, SimpleNamespace(**{**ok, "gait_duty": 1.0})) is None
    assert CentroidalMPC._stance_bound(me, SimpleNamespace(initial_x_vec=None)) is None
