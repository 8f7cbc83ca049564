"""Oracle: restatement of the OSQP algorithm as the reference drives it through CasADi's ``conic``
plugin (``centroidal_mpc.py:20-38`` options, ``:98`` call) on the reference's *sparse* QP.

The solver itself is a third-party dependency that is absent from ``/root/reference`` and not
installable here: OSQP (bundled by CasADi 3.6.7, ``README.md:166``; 0.6.x API).  This file restates
its published algorithm (Stellato et al., "OSQP: an operator splitting solver for quadratic
programs", Math. Prog. Comp. 2020, Alg. 1 + sections 5.1-5.2) with the option values the reference
sets: Ruiz equilibration (``scaling`` passes), per-row rho (equality rows 1e3 rho, free rows
1e-6), over-relaxation alpha = 1.6, sigma = 1e-6, termination test every ``check_termination``
iterations in the scaled space (``scaled_termination``), adaptive rho every
``adaptive_rho_interval`` iterations with the 5x refactor rule, primal + dual warm start, no polish.

**Parity unpinned**: no OSQP binary or recorded OSQP output exists to pin this against; it is
checked by (i) OSQP's own termination criteria evaluated independently, (ii) agreement with the
exact optimum to the accuracy eps implies, (iii) agreement with the C port ``osqp_port.c``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# centroidal_mpc.py:24-35; OSQP defaults for what the reference leaves alone
REFERENCE_OPTS = dict(eps_abs=1e-4, eps_rel=1e-4, max_iter=1000, check_termination=10,
                      adaptive_rho=True, adaptive_rho_interval=25, scaling=5, scaled_termination=True,
                      rho=0.1, sigma=1e-6, alpha=1.6, adaptive_rho_tolerance=5.0)

OSQP_INFTY = 1e30
MIN_SCALING, MAX_SCALING = 1e-4, 1e4
RHO_MIN, RHO_MAX, RHO_TOL, RHO_EQ_OVER_RHO_INEQ = 1e-6, 1e6, 1e-4, 1e3


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


def ruiz_scale(P, q, A, iters):
    """OSQP ``scale_data``: returns (Pbar, qbar, Abar, D, E, c) with Pbar = c D P D, Abar = E A D."""
    n, m = P.shape[0], A.shape[0]
    D, E, c = np.ones(n), np.ones(m), 1.0
    P = sp.csc_matrix(P, dtype=np.float64).copy()
    A = sp.csc_matrix(A, dtype=np.float64).copy()
    q = np.array(q, dtype=np.float64)
    for _ in range(iters):
        ncolP = np.asarray(abs(P).max(axis=0).todense()).ravel()
        ncolA = np.asarray(abs(A).max(axis=0).todense()).ravel()
        nrowA = np.asarray(abs(A).max(axis=1).todense()).ravel()
        Dt = 1.0 / np.sqrt(_limit(np.maximum(ncolP, ncolA)))
        Et = 1.0 / np.sqrt(_limit(nrowA))
        P = sp.diags(Dt) @ P @ sp.diags(Dt)
        A = sp.diags(Et) @ A @ sp.diags(Dt)
        q = Dt * q
        D *= Dt
        E *= Et
        ncolP = np.asarray(abs(P).max(axis=0).todense()).ravel()
        ct = float(_limit(np.array([ncolP.mean()]))[0])
        nq = np.abs(q).max()
        nq = 1.0 if nq < MIN_SCALING else min(nq, MAX_SCALING)
        ct = 1.0 / max(ct, nq)
        P = P * ct
        q = q * ct
        c *= ct
    return sp.csc_matrix(P), q, sp.csc_matrix(A), D, E, c


def rho_vector(l, u, rho):
    free = (l < -OSQP_INFTY * MIN_SCALING) & (u > OSQP_INFTY * MIN_SCALING)
    eq = (u - l) < RHO_TOL
    return np.where(free, RHO_MIN, np.where(eq, RHO_EQ_OVER_RHO_INEQ * rho, rho))


def solve(P, q, A, l, u, x0=None, y0=None, rho0=None, **opts):
    """OSQP on ``min 1/2 x'Px + q'x  s.t. l <= Ax <= u``.  Returns dict(x, y, iters, status, rho,
    nfac, r_prim, r_dual, obj).  ``rho0`` carries the adapted rho of a previous solve."""
    o = dict(REFERENCE_OPTS)
    o.update(opts)
    n, m = P.shape[0], A.shape[0]
    l = np.maximum(np.asarray(l, dtype=np.float64), -OSQP_INFTY)
    u = np.minimum(np.asarray(u, dtype=np.float64), OSQP_INFTY)
    Pb, qb, Ab, D, E, c = ruiz_scale(P, q, A, o["scaling"])
    lb, ub = E * l, E * u
    rho = o["rho"] if rho0 is None else rho0
    sigma, alpha = o["sigma"], o["alpha"]

    def factor(rho):
        rv = rho_vector(lb, ub, rho)
        K = (Pb + sigma * sp.identity(n) + Ab.T @ sp.diags(rv) @ Ab).tocsc()
        return rv, spla.splu(K)

    rv, lu = factor(rho)
    nfac = 1
    x = np.zeros(n) if x0 is None else np.asarray(x0, dtype=np.float64) / D
    y = np.zeros(m) if y0 is None else c * np.asarray(y0, dtype=np.float64) / E
    z = Ab @ x
    status, it = "max_iter", 0
    rp = rd = np.inf
    for it in range(1, o["max_iter"] + 1):
        xt = lu.solve(sigma * x - qb + Ab.T @ (rv * z - y))
        zt = Ab @ xt
        x = alpha * xt + (1 - alpha) * x
        zh = alpha * zt + (1 - alpha) * z
        zn = np.clip(zh + y / rv, lb, ub)
        y = y + rv * (zh - zn)
        z = zn
        check = it % o["check_termination"] == 0
        adapt = o["adaptive_rho"] and o["adaptive_rho_interval"] and it % o["adaptive_rho_interval"] == 0
        if not (check or adapt):
            continue
        Ax, Px, Aty = Ab @ x, Pb @ x, Ab.T @ y
        rp = np.abs(Ax - z).max()
        rd = np.abs(Px + qb + Aty).max()
        np_, nd_ = max(np.abs(Ax).max(), np.abs(z).max()), max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(qb).max())
        if check and rp <= o["eps_abs"] + o["eps_rel"] * np_ and rd <= o["eps_abs"] + o["eps_rel"] * nd_:
            status = "solved"
            break
        if adapt:
            rn = rho * np.sqrt((rp / (np_ + 1e-10)) / (rd / (nd_ + 1e-10) + 1e-10))
            rn = min(max(rn, RHO_MIN), RHO_MAX)
            if rn > rho * o["adaptive_rho_tolerance"] or rn < rho / o["adaptive_rho_tolerance"]:
                rho = rn
                rv, lu = factor(rho)
                nfac += 1
    xs, ys = D * x, E * y / c
    return dict(x=xs, y=ys, iters=it, status=status, rho=rho, nfac=nfac, r_prim=rp, r_dual=rd,
                obj=0.5 * xs @ (P @ xs) + q @ xs)
