"""Oracle: analytic Go2 leg kinematics, the NumPy twin of ``traj::leg_jacobian`` (csrc/cmpc_traj.cuh).

What the reference reads out of Pinocchio (``compute_3x3_foot_Jacobian_world``, go2_robot_data.py:286-300: rows 0-2 of the
LOCAL_WORLD_ALIGNED frame Jacobian of a foot, the three columns of that leg's joints) restated from the published Go2 leg
geometry: hip (abduction) joint about x, thigh and calf joints about y, offsets (0, +-l1, 0), (0, 0, -l2), (0, 0, -l3).
Parity unpinned against Pinocchio (the URDF is not in the reference tree, go2_robot_data.py:24-27); the Jacobian is
checked against finite differences of ``foot_pos_body`` in tests/test_oracle_golden.py.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np

GO2_LINKS = (0.0955, 0.213, 0.213)
SIDE = (1.0, -1.0, 1.0, -1.0)          # FL FR RL RR: left legs carry the abduction offset towards +y


def _rx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _ry(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def foot_pos_body(q3, side, links=GO2_LINKS):
    """Foot relative to its hip in the body frame, by composing the joint rotations link by link."""
    l1, l2, l3 = links
    p = np.array([0.0, 0.0, -l3])
    p = _ry(q3[2]) @ p + np.array([0.0, 0.0, -l2])
    p = _ry(q3[1]) @ p + np.array([0.0, side * l1, 0.0])
    return _rx(q3[0]) @ p


def jacobian_world(q3, R_world_to_body, side, links=GO2_LINKS):
    """3 x 3 world-aligned translational Jacobian of one foot over (hip, thigh, calf): each column is the joint axis
    (in the body frame) crossed with the lever from the joint to the foot, rotated to the world."""
    l1, l2, l3 = links
    R1 = _rx(q3[0])
    R12 = R1 @ _ry(q3[1])
    o_thigh = R1 @ np.array([0.0, side * l1, 0.0])
    o_calf = o_thigh + R12 @ np.array([0.0, 0.0, -l2])
    p = foot_pos_body(q3, side, links)
    ax = [np.array([1.0, 0, 0]), R1 @ np.array([0, 1.0, 0]), R12 @ np.array([0, 1.0, 0])]
    org = [np.zeros(3), o_thigh, o_calf]
    Jb = np.stack([np.cross(a, p - o) for a, o in zip(ax, org)], axis=1)
    return np.asarray(R_world_to_body).T @ Jb
