"""Oracle: condensed form of the reference QP (not present in the reference --
it follows algebraically from the equality rows of ``centroidal_mpc.py:287-303``).

    X = A_qp x0 + B_qp U + G,   H = 2 (B_qp' L B_qp + K),   g = 2 B_qp' L (A_qp x0 + G - x_ref)
    U = [u_0; ...; u_{N-1}]  (12N),  u_k = [f_FL; f_FR; f_RL; f_RR]

Constraint rows keep the reference ordering: 12N box rows on U (``centroidal_mpc.py:122-176``) then
16N friction rows ``16k + 4 leg + face`` (``:324-359``, bounds ``:264-279``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np

from .sparse_qp import COST_Q, COST_R, MU, FZ_MIN


def prediction_matrices(Ad, Bd, gd):
    """A_qp (12N,12), B_qp (12N,12N) block lower-triangular, G (12N,)."""
    Ad = np.asarray(Ad, dtype=np.float64)
    Bd = np.asarray(Bd, dtype=np.float64)
    gd = np.asarray(gd, dtype=np.float64).reshape(12)
    N = Bd.shape[0]
    Aqp = np.zeros((12 * N, 12))
    Bqp = np.zeros((12 * N, 12 * N))
    G = np.zeros(12 * N)
    P = np.eye(12)
    acc = np.zeros(12)
    for i in range(N):
        acc = Ad @ acc + gd            # G_i = sum_{j<=i} Ad^j gd
        G[12 * i:12 * i + 12] = acc
        P = Ad @ P                     # Ad^(i+1)
        Aqp[12 * i:12 * i + 12] = P
    for a in range(N):
        blk = Bd[a].copy()
        for i in range(a, N):
            Bqp[12 * i:12 * i + 12, 12 * a:12 * a + 12] = blk
            blk = Ad @ blk
    return Aqp, Bqp, G


def build(Ad, Bd, gd, x0, x_ref, contact, Q=COST_Q, R=COST_R, mu=MU, fz_min=FZ_MIN):
    """Dense condensed QP for one robot: dict(H, g, A, l, u, Aqp, Bqp, G, c0)."""
    x0 = np.asarray(x0, dtype=np.float64).reshape(12)
    x_ref = np.asarray(x_ref, dtype=np.float64)
    contact = np.asarray(contact)
    N = np.asarray(Bd).shape[0]
    n = 12 * N
    Aqp, Bqp, G = prediction_matrices(Ad, Bd, gd)
    Lq = np.tile(Q, N)
    Kr = np.tile(R, N)
    e0 = Aqp @ x0 + G - x_ref.reshape(-1, order="F")
    H = 2.0 * (Bqp.T @ (Lq[:, None] * Bqp) + np.diag(Kr))
    H = 0.5 * (H + H.T)
    g = 2.0 * (Bqp.T @ (Lq * e0))
    # constant that makes the condensed objective equal the sparse-form objective
    # sum (x-xr)'Q(x-xr) + u'Ru - xr'Q xr  ->  sparse objective drops xr'Q xr
    xr = x_ref.reshape(-1, order="F")
    c0 = e0 @ (Lq * e0) - xr @ (Lq * xr)
    A, l, u = constraints(contact, mu, fz_min)
    return dict(H=H, g=g, A=A, l=l, u=u, Aqp=Aqp, Bqp=Bqp, G=G, c0=c0, N=N, x0=x0)


def constraints(contact, mu=MU, fz_min=FZ_MIN):
    """A = [I_12N ; F] (28N x 12N), l, u in the reference's row order."""
    contact = np.asarray(contact)
    N = contact.shape[1]
    n = 12 * N
    F = np.zeros((16 * N, n))
    lb = np.full(n, -np.inf)
    ub = np.full(n, np.inf)
    uf = np.full(16 * N, np.inf)
    for k in range(N):
        for leg in range(4):
            j = 12 * k + 3 * leg
            r = 16 * k + 4 * leg
            F[r + 0, j + 0], F[r + 0, j + 2] = 1.0, -mu
            F[r + 1, j + 0], F[r + 1, j + 2] = -1.0, -mu
            F[r + 2, j + 1], F[r + 2, j + 2] = 1.0, -mu
            F[r + 3, j + 1], F[r + 3, j + 2] = -1.0, -mu
            if contact[leg, k] == 1:
                lb[j + 2] = fz_min
                uf[r:r + 4] = 0.0
            else:
                lb[j:j + 3] = 0.0
                ub[j:j + 3] = 0.0
    A = np.vstack([np.eye(n), F])
    l = np.concatenate([lb, np.full(16 * N, -np.inf)])
    u = np.concatenate([ub, uf])
    return A, l, u


def rollout(cq, U):
    """States x_1..x_N (12N,) implied by forces U (12N,)."""
    return cq["Aqp"] @ cq["x0"] + cq["Bqp"] @ U + cq["G"]


def lift(cq, U, y=None, Ad=None, Bd=None, x_ref=None, Q=COST_Q):
    """Condensed (U, y) -> reference variables ``w = [X; U]`` and duals ``(lam_x, lam_a)``.

    y (28N,) are the condensed duals in row order [box(12N); friction(16N)].  The equality duals
    (co-states) follow the backward recursion of SURVEY.md Appendix B.
    """
    N = cq["N"]
    X = rollout(cq, U)
    w = np.concatenate([X, U])
    if y is None:
        return w
    lam_x = np.concatenate([np.zeros(12 * N), y[:12 * N]])
    xr = np.asarray(x_ref).reshape(-1, order="F")
    nu = np.zeros((N, 12))
    for k in range(N - 1, -1, -1):
        xk = X[12 * k:12 * k + 12]
        v = -2.0 * Q * (xk - xr[12 * k:12 * k + 12])
        if k < N - 1:
            v = v + np.asarray(Ad).T @ nu[k + 1]
        nu[k] = v
    lam_a = np.concatenate([nu.reshape(-1), y[12 * N:]])
    return w, lam_x, lam_a


def objective(cq, U):
    return 0.5 * U @ (cq["H"] @ U) + cq["g"] @ U


def kkt_residuals(cq, U, y):
    """Independent optimality certificate for (U, y): stationarity, primal infeasibility,
    dual-sign violation and complementarity, all in inf-norm (absolute)."""
    A, l, u = cq["A"], cq["l"], cq["u"]
    stat = np.abs(cq["H"] @ U + cq["g"] + A.T @ y).max()
    z = A @ U
    prim = max(np.maximum(l - z, 0).max(), np.maximum(z - u, 0).max())
    # y>0 only where z=u, y<0 only where z=l
    du = np.where(np.isfinite(u), 0.0, np.maximum(y, 0)).max()
    dl = np.where(np.isfinite(l), 0.0, np.maximum(-y, 0)).max()
    gap_u = np.where(np.isfinite(u), np.maximum(y, 0) * np.abs(u - z), 0.0)
    gap_l = np.where(np.isfinite(l), np.maximum(-y, 0) * np.abs(z - l), 0.0)
    gap_u = np.nan_to_num(gap_u, nan=0.0)
    gap_l = np.nan_to_num(gap_l, nan=0.0)
    comp = max(gap_u.max(), gap_l.max())
    return dict(stat=stat, prim=prim, dual=max(du, dl), comp=comp)
