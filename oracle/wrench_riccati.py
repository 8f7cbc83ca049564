"""Oracle: wrench-space projected Riccati + primal-dual active set, the NumPy twin of the device kernel
``csrc/cmpc_wrench.cuh``.

The reference hands the QP of ``centroidal_mpc.py:69-120`` to OSQP.  The device solves the same QP by a
primal-dual active-set whose equality-constrained sub-problems are solved by a Riccati sweep over the
horizon.  Two structural facts make that sweep uniform across robots, stance patterns and working sets:

* every foot force enters the dynamics of ``com_trajectory.py:234-262`` through the 6-dimensional wrench
  ``w = sum_j [I; W_j] f_j`` (``W_j = I_world^-1 [r_j]x``):  ``B_d[k] = Bbar [C_1 .. C_4]`` with the constant
  12 x 6 matrix ``Bbar = [[h/m I, 0], [0, h Rz'], [dt/m I, 0], [0, dt I]]``;
* every inequality row (``fz >= fz_min`` ``centroidal_mpc.py:163-170``, friction faces ``:324-359``) touches a
  single foot at a single step, so a working set is eliminated foot by foot: ``f_j = fhat_j - Pi_j C_j' mu``
  with ``Pi_j = Z (Z'RZ)^-1 Z'`` (``R^-1`` for a free foot, 0 for a swing or fully pinned foot).

With ``Lam_k = sum_j C_j Pi_j C_j'`` the stage minimisation needs only 6 x 6 factorizations:

    Gbar = Bbar' P Bbar = L L',  Y = L^-1 Bbar' P A,  Nn = I + L' Lam L,
    P <- Q + A'PA - Y'(I - Nn^-1)Y,   mu_k = L Nn^-1 (Y x_k + L^-1 Bbar' q)

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np

from .sparse_qp import COST_Q, COST_R, MU, FZ_MIN

GRAV = 9.81

# active code of one foot-step: bit 0 fz pinned at fz_min; bits 1-2 fx: 0 free, 1 fx = +mu fz, 2 fx = -mu fz;
# bits 3-4 fy likewise.  SWING marks a foot that is not in stance (force pinned to zero, centroidal_mpc.py:150-161).
SWING = 255


def code_of(az, ax, ay):
    return int(az) | (int(ax) << 1) | (int(ay) << 3)


def foot_projection(code, R3, mu, fz_min):
    """(Pi (3,3), fhat (3,)) of one foot-step for the working-set code."""
    if code == SWING:
        return np.zeros((3, 3)), np.zeros(3)
    az, ax, ay = code & 1, (code >> 1) & 3, (code >> 3) & 3
    sx = 0.0 if ax == 0 else (mu if ax == 1 else -mu)
    sy = 0.0 if ay == 0 else (mu if ay == 1 else -mu)
    Pi = np.zeros((3, 3))
    fhat = np.zeros(3)
    if ax == 0:
        Pi[0, 0] = 1.0 / R3[0]
    if ay == 0:
        Pi[1, 1] = 1.0 / R3[1]
    if az:
        fhat[:] = (sx * fz_min, sy * fz_min, fz_min)
    else:
        z = np.array([sx, sy, 1.0])
        Pi += np.outer(z, z) / (z @ (R3 * z))
    return Pi, fhat


def foot_multipliers(code, d, mu):
    """lam (5,) >= 0 wanted, rows (fz_min, +fx, -fx, +fy, -fy), from 2 d + A_act' lam = 0 (d = R f + C' mu)."""
    lam = np.zeros(5)
    if code == SWING:
        return lam
    az, ax, ay = code & 1, (code >> 1) & 3, (code >> 3) & 3
    # rows: r0 = (0,0,-1); r1 = (1,0,-mu); r2 = (-1,0,-mu); r3 = (0,1,-mu); r4 = (0,-1,-mu)
    lx = ly = 0.0
    if ax == 1:
        lx = -2.0 * d[0]; lam[1] = lx
    elif ax == 2:
        lx = 2.0 * d[0]; lam[2] = lx
    if ay == 1:
        ly = -2.0 * d[1]; lam[3] = ly
    elif ay == 2:
        ly = 2.0 * d[1]; lam[4] = ly
    if az:
        lam[0] = 2.0 * d[2] - mu * (lx + ly)
    return lam


def foot_viol(f, mu, fz_min):
    return np.array([fz_min - f[2], f[0] - mu * f[2], -f[0] - mu * f[2], f[1] - mu * f[2], -f[1] - mu * f[2]])


class Robot:
    """Raw inputs of one robot (records.py layout) with the wrench-space matrices."""

    def __init__(self, x0, x_ref, r_foot, I_world, mass, dt, contact, Q=COST_Q, R=COST_R, mu=MU, fz_min=FZ_MIN):
        self.N = N = x_ref.shape[1]
        self.x0, self.xr = np.asarray(x0, float), np.asarray(x_ref, float)
        self.contact = np.asarray(contact)
        self.Q, self.R, self.mu, self.fz_min, self.dt = np.asarray(Q, float), np.asarray(R, float), mu, fz_min, dt
        yaw = np.average(self.xr[5, :])                       # com_trajectory.py:226
        c, s = np.cos(yaw), np.sin(yaw)
        Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        h = dt * dt / 2
        self.A = np.eye(12)
        self.A[0:3, 6:9] = dt * np.eye(3)
        self.A[3:6, 9:12] = dt * Rz.T
        Bb = np.zeros((12, 6))
        Bb[0:3, 0:3] = h / mass * np.eye(3)
        Bb[3:6, 3:6] = h * Rz.T
        Bb[6:9, 0:3] = dt / mass * np.eye(3)
        Bb[9:12, 3:6] = dt * np.eye(3)
        self.Bb = Bb
        self.g = np.zeros(12); self.g[2] = -GRAV * h; self.g[8] = -GRAV * dt
        Iinv = np.linalg.inv(I_world)
        self.C = np.zeros((N, 4, 6, 3))
        for k in range(N):
            for j in range(4):
                r = r_foot[j, :, k]
                sk = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
                self.C[k, j, 0:3] = np.eye(3)
                self.C[k, j, 3:6] = Iinv @ sk

    def unconstrained_codes(self):
        return np.where(self.contact.T == 1, 0, SWING).astype(np.int32)          # (N,4)


def sweep(rb, codes):
    """One equality-constrained solve for the working set ``codes`` (N,4).  Returns forces F (N,4,3), states
    X (N,12) = x_1..x_N, wrench co-states mu (N,6), multipliers lam (N,4,5) and the smallest pivot seen."""
    N, A, Bb, Q = rb.N, rb.A, rb.Bb, rb.Q
    Pi = np.zeros((N, 4, 3, 3)); fh = np.zeros((N, 4, 3))
    Lam = np.zeros((N, 6, 6)); what = np.zeros((N, 6))
    for k in range(N):
        for j in range(4):
            Pi[k, j], fh[k, j] = foot_projection(int(codes[k, j]), rb.R[3 * j:3 * j + 3], rb.mu, rb.fz_min)
            C = rb.C[k, j]
            Lam[k] += C @ Pi[k, j] @ C.T
            what[k] += C @ fh[k, j]
    P = np.diag(Q).copy()
    p = -Q * rb.xr[:, N - 1]
    Kb = np.zeros((N, 6, 12)); kb = np.zeros((N, 6))
    pmin = np.inf
    for k in range(N - 1, -1, -1):
        PB = P @ Bb
        Gb = Bb.T @ PB
        L = np.linalg.cholesky(Gb)
        pmin = min(pmin, np.diag(L).min())
        S = PB.T @ A
        Y = np.linalg.solve(L, S)
        Nn = np.eye(6) + L.T @ Lam[k] @ L
        Ln = np.linalg.cholesky(Nn)
        ghat = rb.g + Bb @ what[k]
        q = P @ ghat + p
        yq = np.linalg.solve(L, Bb.T @ q)
        NiY = np.linalg.solve(Ln.T, np.linalg.solve(Ln, Y))
        Niq = np.linalg.solve(Ln.T, np.linalg.solve(Ln, yq))
        Kb[k] = L @ NiY
        kb[k] = L @ Niq
        if k > 0:
            P = np.diag(Q) + A.T @ P @ A - Y.T @ (Y - NiY)
            P = 0.5 * (P + P.T)
            p = -Q * rb.xr[:, k - 1] + A.T @ q - Y.T @ (yq - Niq)
    X = np.zeros((N, 12)); F = np.zeros((N, 4, 3)); mus = np.zeros((N, 6)); lam = np.zeros((N, 4, 5))
    x = rb.x0.copy()
    for k in range(N):
        m = Kb[k] @ x + kb[k]
        mus[k] = m
        w = np.zeros(6)
        for j in range(4):
            C = rb.C[k, j]
            F[k, j] = fh[k, j] - Pi[k, j] @ (C.T @ m)
            w += C @ F[k, j]
            if codes[k, j] != SWING:
                d = rb.R[3 * j:3 * j + 3] * F[k, j] + C.T @ m
                lam[k, j] = foot_multipliers(int(codes[k, j]), d, rb.mu)
        x = A @ x + rb.g + Bb @ w
        X[k] = x
    return F, X, mus, lam, pmin


def next_codes(rb, codes, F, lam, tol=1e-10):
    """Primal-dual active-set update, foot by foot (same rule as solve_active_set_fast in cmpc_fast.cuh):
    row t joins the working set iff lam_t + viol_t > tol; of two opposite faces only the larger one."""
    new = codes.copy()
    for k in range(rb.N):
        for j in range(4):
            if codes[k, j] == SWING:
                continue
            s = lam[k, j] + foot_viol(F[k, j], rb.mu, rb.fz_min)
            az = s[0] > tol
            ax = 1 if (s[1] > tol and s[1] >= s[2]) else (2 if (s[2] > tol and s[2] > s[1]) else 0)
            ay = 1 if (s[3] > tol and s[3] >= s[4]) else (2 if (s[4] > tol and s[4] > s[3]) else 0)
            new[k, j] = code_of(az, ax, ay)
    return new


def solve(rb, max_iter=16, codes=None):
    """PDAS from the unconstrained working set (or a warm one).  dict(U (12N,), X, lam, iters, ok, codes)."""
    codes = rb.unconstrained_codes() if codes is None else codes.copy()
    seen = []
    ok = False
    it = 0
    for it in range(1, max_iter + 1):
        F, X, mus, lam, pmin = sweep(rb, codes)
        new = next_codes(rb, codes, F, lam)
        if (new == codes).all():
            ok = True
            break
        if any((new == s).all() for s in seen):
            break                                   # cycle
        seen.append(codes)
        codes = new
    U = F.reshape(rb.N, 12).reshape(-1)
    return dict(U=U, X=X, lam=lam, iters=it, ok=ok, codes=codes, F=F, mu=mus)


def duals_condensed(rb, res):
    """y (28N,) in the reference's row order from the foot multipliers (stance rows only; swing box rows are
    filled by the caller from stationarity)."""
    N = rb.N
    y = np.zeros(28 * N)
    for k in range(N):
        for j in range(4):
            l = res["lam"][k, j]
            y[12 * k + 3 * j + 2] = -l[0]
            y[12 * N + 16 * k + 4 * j:12 * N + 16 * k + 4 * j + 4] = l[1:5]
    return y
