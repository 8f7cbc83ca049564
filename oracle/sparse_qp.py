"""Oracle: the reference's *sparse* (non-condensed) MPC QP, built exactly as
``convex_mpc/centroidal_mpc.py`` builds it.

    w = [x_1..x_N ; u_0..u_{N-1}]  (24N)          centroidal_mpc.py:44, test_MPC.py:189-192
    min 1/2 w'Hw + g'w   s.t.  lba <= A w <= uba,  lbx <= w <= ubx

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np
import scipy.sparse as sp

# centroidal_mpc.py:12-17
COST_Q = np.array([1, 1, 50, 10, 20, 1, 2, 2, 1, 1, 1, 1], dtype=np.float64)
COST_R = np.full(12, 1e-5)
MU = 0.8
NX = 12
NU = 12
FZ_MIN = 10.0   # centroidal_mpc.py:127


def hessian(N, Q=COST_Q, R=COST_R):
    """blkdiag(2Q x N, 2R x N) -- centroidal_mpc.py:183-201."""
    d = np.concatenate([np.tile(2 * Q, N), np.tile(2 * R, N)])
    return sp.diags(d).tocsc()


def friction_matrix(N, mu=MU):
    """(16N x 24N): row 16k+4leg+face, faces (+fx,-fx,+fy,-fy) - mu fz  (centroidal_mpc.py:324-359)."""
    rows, cols, vals = [], [], []
    r = 0
    for k in range(N):
        base = N * NX + k * NU
        for leg in range(4):
            fx, fy, fz = base + 3 * leg, base + 3 * leg + 1, base + 3 * leg + 2
            for (c, sgn) in ((fx, 1.0), (fx, -1.0), (fy, 1.0), (fy, -1.0)):
                rows += [r, r]
                cols += [c, fz]
                vals += [sgn, -mu]
                r += 1
    return sp.csc_matrix((vals, (rows, cols)), shape=(r, N * (NX + NU)))


def dynamics_matrix(Ad, Bd):
    """A_eq (12N x 24N): row block k = [+I at x_{k+1}, -Ad at x_k (k>=1), -Bd[k] at u_k]
    (centroidal_mpc.py:287-303, 305-321)."""
    N = Bd.shape[0]
    I = sp.identity(N * NX, format="csc")
    shift = sp.kron(sp.diags([np.ones(N - 1)], [-1]), sp.identity(NX))      # :54
    big_minus_Ad = sp.block_diag([-np.asarray(Ad)] * N)
    big_minus_Bd = sp.block_diag([-np.asarray(Bd[k]) for k in range(N)])
    return sp.hstack([I + shift @ big_minus_Ad, big_minus_Bd]).tocsc()


def structural_pattern(N):
    """Boolean (28N x 24N) pattern of the constraint matrix as the reference creates it (``self.A_sp``,
    centroidal_mpc.py:203-209): identity + dense -Ad blocks on the sub-diagonal, dense -Bd blocks, friction rows."""
    pat = np.zeros((28 * N, 24 * N), dtype=bool)
    for k in range(N):
        pat[12 * k:12 * k + 12, 12 * k:12 * k + 12] |= np.eye(12, dtype=bool)
        if k >= 1:
            pat[12 * k:12 * k + 12, 12 * (k - 1):12 * k] = True
        pat[12 * k:12 * k + 12, 12 * N + 12 * k:12 * N + 12 * k + 12] = True
    pat[12 * N:, :] = friction_matrix(N).toarray() != 0
    return pat


def build(Ad, Bd, gd, x0, x_ref, contact, Q=COST_Q, R=COST_R, mu=MU, fz_min=FZ_MIN):
    """Return dict(H, g, A, lba, uba, lbx, ubx) for one robot, reference ordering.

    Ad (12,12), Bd (N,12,12), gd (12,) or (12,1), x0 (12,), x_ref (12,N), contact (4,N) 0/1.
    """
    Ad = np.asarray(Ad, dtype=np.float64)
    Bd = np.asarray(Bd, dtype=np.float64)
    gd = np.asarray(gd, dtype=np.float64).reshape(12)
    x0 = np.asarray(x0, dtype=np.float64).reshape(12)
    x_ref = np.asarray(x_ref, dtype=np.float64)
    contact = np.asarray(contact)
    N = Bd.shape[0]
    nv = N * (NX + NU)

    H = hessian(N, Q, R)
    # g = [vec_colmajor(-2 Q x_ref); 0]  (centroidal_mpc.py:248-253)
    g = np.concatenate([(-2.0 * (Q[:, None] * x_ref)).reshape(-1, order="F"), np.zeros(N * NU)])
    A = sp.vstack([dynamics_matrix(Ad, Bd), friction_matrix(N, mu)]).tocsc()
    # beq (centroidal_mpc.py:257-261); the 12-term products are accumulated in ascending column order with separately
    # rounded multiply and add, the order of CasADi's dense mtimes (column-by-column axpy)
    first = np.zeros(12)
    for r in range(12):
        first = first + Ad[:, r] * x0[r]
    beq = np.concatenate([first + gd] + [gd] * (N - 1))
    # friction bounds: (-inf, 0] stance, (-inf, +inf) swing; row order k-major, leg, face (:264-279)
    u_ineq = np.full(16 * N, np.inf)
    for k in range(N):
        for leg in range(4):
            if contact[leg, k] == 1:
                u_ineq[16 * k + 4 * leg:16 * k + 4 * leg + 4] = 0.0
    l_ineq = np.full(16 * N, -np.inf)
    lba = np.concatenate([beq, l_ineq])
    uba = np.concatenate([beq, u_ineq])
    # variable bounds (centroidal_mpc.py:122-176)
    lbx = np.full(nv, -np.inf)
    ubx = np.full(nv, np.inf)
    for k in range(N):
        for leg in range(4):
            j = N * NX + k * NU + 3 * leg
            if contact[leg, k] == 1:
                lbx[j + 2] = max(lbx[j + 2], fz_min)
            else:
                lbx[j:j + 3] = 0.0
                ubx[j:j + 3] = 0.0
    return dict(H=H, g=g, A=A, lba=lba, uba=uba, lbx=lbx, ubx=ubx, N=N)


def split_solution(w, N):
    """test_MPC.py:189-192 -- X (12,N), U (12,N), both column-major slices of w."""
    w = np.asarray(w).reshape(-1)
    X = w[:12 * N].reshape((12, N), order="F")
    U = w[12 * N:].reshape((12, N), order="F")
    return X, U


def objective(qp, w):
    return 0.5 * w @ (qp["H"] @ w) + qp["g"] @ w


def as_osqp_form(qp):
    """What the CasADi osqp plugin hands OSQP: A_osqp = [I; A], l = [lbx; lba], u = [ubx; uba],
    +-inf clipped to +-1e30 is left to the solver ([recall])."""
    n = qp["H"].shape[0]
    A = sp.vstack([sp.identity(n, format="csc"), qp["A"]]).tocsc()
    l = np.concatenate([qp["lbx"], qp["lba"]])
    u = np.concatenate([qp["ubx"], qp["uba"]])
    return qp["H"], qp["g"], A, l, u
