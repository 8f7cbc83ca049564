"""Oracle: high-accuracy solution of the condensed QP with an independent KKT certificate.

Used as the L2 parity target ("forces vs the exact optimum").  Method: eliminate
variables fixed by equal bounds (swing legs, ``centroidal_mpc.py:150-161``), run a dense ADMM to
get an active-set guess, then iterate a primal-dual active-set refinement (solve the equality-
constrained KKT system on the working set; add violated rows, drop rows with wrong-signed
multipliers) until the KKT conditions hold to ~1e-9.  The certificate is computed by
``condensed_qp.kkt_residuals`` from (U, y) alone, so it does not depend on how they were found.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np
import scipy.linalg as sla


def _reduce(H, g, A, l, u):
    """Split variables into fixed (box row with l==u) and free; return reduced problem pieces."""
    n = H.shape[0]
    fixed = np.isfinite(l[:n]) & (l[:n] == u[:n])
    free = ~fixed
    xf = np.where(fixed, l[:n], 0.0)
    Hr = H[np.ix_(free, free)]
    gr = g[free] + H[np.ix_(free, fixed)] @ xf[fixed]
    # rows that still matter: finite bound somewhere and touching a free variable, not the fixed box rows
    rows = np.ones(A.shape[0], dtype=bool)
    rows[:n] = free & (np.isfinite(l[:n]) | np.isfinite(u[:n]))
    rows[n:] = np.isfinite(l[n:]) | np.isfinite(u[n:])
    Ar = A[np.ix_(rows, free)]
    shift = A[np.ix_(rows, fixed)] @ xf[fixed]
    return free, fixed, xf, rows, Hr, gr, Ar, l[rows] - shift, u[rows] - shift


def admm_dense(H, g, A, l, u, rho=1e-4, sigma=1e-6, alpha=1.6, iters=2000, eps=1e-9,
               x=None, y=None):
    """Plain OSQP-style ADMM on a dense problem with a fixed scalar rho (no scaling)."""
    n, m = H.shape[0], A.shape[0]
    K = H + sigma * np.eye(n) + rho * (A.T @ A)
    cf = sla.cho_factor(K)
    x = np.zeros(n) if x is None else x.copy()
    y = np.zeros(m) if y is None else y.copy()
    z = np.clip(A @ x, l, u)
    it = 0
    for it in range(1, iters + 1):
        xt = sla.cho_solve(cf, sigma * x - g + A.T @ (rho * z - y))
        zt = A @ xt
        x = alpha * xt + (1 - alpha) * x
        zh = alpha * zt + (1 - alpha) * z
        zn = np.clip(zh + y / rho, l, u)
        y = y + rho * (zh - zn)
        z = zn
        if it % 10 == 0:
            Ax = A @ x
            rp = np.abs(Ax - z).max()
            rd = np.abs(H @ x + g + A.T @ y).max()
            if rp <= eps * (1 + max(np.abs(Ax).max(), np.abs(z).max())) and \
               rd <= eps * (1 + max(np.abs(H @ x).max(), np.abs(A.T @ y).max(), np.abs(g).max())):
                break
    return x, y, z, it


def _kkt_solve(H, g, A, b, act):
    """min 1/2 x'Hx + g'x  s.t.  A[act] x = b[act]  ->  (x, lambda)."""
    n = H.shape[0]
    k = int(act.sum())
    if k == 0:
        return sla.solve(H, -g, assume_a="pos"), np.zeros(0)
    Aa = A[act]
    cf = sla.cho_factor(H)
    HiAt = sla.cho_solve(cf, Aa.T)
    Hig = sla.cho_solve(cf, g)
    S = Aa @ HiAt
    S = 0.5 * (S + S.T) + 1e-14 * np.eye(k) * np.trace(S) / max(k, 1)
    lam = np.linalg.lstsq(S, -(b[act] + Aa @ Hig), rcond=1e-13)[0]
    x = -(Hig + HiAt @ lam)
    return x, lam


def solve_exact(H, g, A, l, u, max_refine=60, tol=1e-9):
    """Return dict(U, y, iters, refine, ok).  ``y`` are duals for *all* rows of A (sign: y>0 at
    upper bounds, y<0 at lower bounds; fixed variables get the multiplier that closes stationarity)."""
    n = H.shape[0]
    free, fixed, xf, rows, Hr, gr, Ar, lr, ur = _reduce(H, g, A, l, u)
    m = Ar.shape[0]
    if Hr.shape[0] == 0:
        # every variable is fixed (all legs in swing): nothing to optimise
        U = xf.copy()
        yfull = np.zeros(A.shape[0])
        yfull[:n] = -(H @ U + g)
        return dict(U=U, y=yfull, iters=0, refine=0, ok=True)
    x, y, z, it = admm_dense(Hr, gr, Ar, lr, ur, rho=1e-4, iters=3000, eps=1e-8)
    Ax = Ar @ x
    scale = 1.0 + np.abs(Ax)
    up = np.isfinite(ur) & ((y > 1e-9) | (Ax >= ur - 1e-7 * scale)) & (y >= -1e-12)
    lo = np.isfinite(lr) & ((y < -1e-9) | (Ax <= lr + 1e-7 * scale)) & (y <= 1e-12)
    ok = False
    refine = 0
    lam_full = np.zeros(m)
    for refine in range(1, max_refine + 1):
        act = up | lo
        b = np.where(up, ur, np.where(lo, lr, 0.0))
        xs, lam = _kkt_solve(Hr, gr, Ar, b, act)
        lam_full = np.zeros(m)
        lam_full[act] = lam
        Ax = Ar @ xs
        viol_u = np.isfinite(ur) & (Ax > ur + tol * (1 + np.abs(ur))) & ~act
        viol_l = np.isfinite(lr) & (Ax < lr - tol * (1 + np.abs(lr))) & ~act
        bad_u = up & (lam_full < -tol)
        bad_l = lo & (lam_full > tol)
        if not (viol_u.any() or viol_l.any() or bad_u.any() or bad_l.any()):
            ok = True
            x = xs
            break
        # drop the worst wrong-signed row first; otherwise add the most violated row
        if bad_u.any() or bad_l.any():
            score = np.where(bad_u, -lam_full, 0.0) + np.where(bad_l, lam_full, 0.0)
            j = int(np.argmax(score))
            up[j] = False
            lo[j] = False
        else:
            vu = np.where(viol_u, Ax - ur, 0.0)
            vl = np.where(viol_l, lr - Ax, 0.0)
            if vu.max() >= vl.max():
                up[int(np.argmax(vu))] = True
            else:
                lo[int(np.argmax(vl))] = True
        x = xs
    U = xf.copy()
    U[free] = x
    yfull = np.zeros(A.shape[0])
    yfull[np.flatnonzero(rows)] = lam_full
    # multipliers of the fixed variables close the stationarity equation exactly
    r = H @ U + g + A.T @ yfull
    yfull[:n][fixed] = -r[fixed]
    return dict(U=U, y=yfull, iters=it, refine=refine, ok=ok)
