"""Oracle: yaw-rotated single-rigid-body dynamics and its ZOH discretisation.

Restates reference ``convex_mpc/com_trajectory.py:15-25`` (reference stack),
``:213-219`` (skew), ``:221-270`` (continuous model), ``:272-286`` (discretisation).
Two routes are kept: the *literal* one (SciPy ``cont2discrete`` + 50-point trapezoid of ``expm``)
and the *closed form* that follows from ``A_c @ A_c == 0``.  Tests check they agree to ~1e-17.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np

GRAVITY = 9.81  # com_trajectory.py:268


def skew(v):
    """com_trajectory.py:213-219."""
    x, y, z = v
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def x_ref_vec(pos, rpy, vel, omega):
    """(12, N) reference, rows [p; rpy; v; omega] (com_trajectory.py:15-25)."""
    refs = [np.asarray(r, dtype=np.float64) for r in (pos, rpy, vel, omega)]
    n = min(r.shape[1] for r in refs)
    return np.vstack([r[:, :n] for r in refs])


def yaw_average(x_ref):
    """com_trajectory.py:226 -- plain mean of the yaw reference row."""
    return np.average(np.asarray(x_ref)[5, :])


def continuous_dynamics(mass, I_world, yaw_avg, r_foot):
    """A_c (12,12), B_c (N,12,12), g_c (12,) -- com_trajectory.py:221-270.

    ``r_foot``: (4, 3, N) lever arms CoM->foot in the world frame, legs FL, FR, RL, RR.
    """
    r_foot = np.asarray(r_foot, dtype=np.float64)
    N = r_foot.shape[2]
    c, s = np.cos(yaw_avg), np.sin(yaw_avg)
    Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    Ac = np.zeros((12, 12))
    Ac[0:3, 6:9] = np.eye(3)          # p_dot = v
    Ac[3:6, 9:12] = Rz.T              # rpy_dot = Rz(yaw_avg)^T omega
    I_inv = np.linalg.inv(np.asarray(I_world, dtype=np.float64))   # :255
    Bc = np.zeros((N, 12, 12))
    for i in range(N):
        for leg in range(4):
            Bc[i, 6:9, 3 * leg:3 * leg + 3] = (1 / mass) * np.eye(3)      # :260
            Bc[i, 9:12, 3 * leg:3 * leg + 3] = I_inv @ skew(r_foot[leg, :, i])   # :261
    gc = np.zeros(12)
    gc[8] = -GRAVITY
    return Ac, Bc, gc


def discrete_dynamics_literal(Ac, Bc, gc, dt):
    """The reference's own route (com_trajectory.py:272-286): SciPy ZOH + trapezoid of expm."""
    from scipy.signal import cont2discrete
    from scipy.linalg import expm
    N = Bc.shape[0]
    Bd = np.zeros((N, 12, 12))
    Ad = None
    for i in range(N):
        Ad, Bd[i], *_ = cont2discrete((Ac, Bc[i], np.eye(12), np.zeros((12, 12))), dt, method="zoh")
    tau = np.linspace(0, dt, 50)
    terms = np.stack([expm(Ac * t) @ gc for t in tau], axis=1)
    trapz = getattr(np, "trapezoid", None) or np.trapz   # same rule; np.trapz is the reference's spelling
    gd = trapz(terms, tau, axis=1)
    return Ad, Bd, gd.reshape(-1, 1)


def discrete_dynamics_closed(Ac, Bc, gc, dt):
    """Closed form: A_c^2 = 0  =>  A_d = I + dt A_c,  B_d = (dt I + dt^2/2 A_c) B_c,
    g_d = (dt I + dt^2/2 A_c) g_c  (the trapezoid rule is exact for the affine integrand)."""
    Ad = np.eye(12) + dt * Ac
    M = dt * np.eye(12) + (dt * dt / 2) * Ac
    Bd = np.einsum("ij,njk->nik", M, Bc)
    gd = M @ gc
    return Ad, Bd, gd.reshape(-1, 1)
