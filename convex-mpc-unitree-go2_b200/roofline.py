"""Algorithmic work of the fused solve kernel per QP, used for the roofline
fraction that ``bench.py`` reports.  "Algorithmic" = the flops the mathematics needs for the route
each QP actually took (its number of free variables n, active-set solves, ADMM iterations and
factorisations come from the kernel's own per-QP statistics) -- not the instructions executed.

    build        0.4e6 * (n/192)^2      closed-form condensed H, g   
    cholesky     n^3 / 3
    tri-solves   2 n^2                  forward + backward substitution for the unconstrained minimiser
    inverse      n^3 / 3                W = L^-1, only when an active-set / ADMM phase runs
    active set   per working-set solve with k active rows:  k^2 n + k^3/3 + 2 n^2
    ADMM         per iteration 2 n^2 + 12 (n + m), m = 5n/3;  per re-factorisation  build + 2 n^3 / 3
    epilogue     roll-out + co-states + residuals: 2 * (2*144*N + 2*12*n) + 4 n

Bytes: every QP reads 3 280 B of inputs and writes u, y, X, nu, status, iters, stats
(8*(12N + 28N + 12N + 12N) + 8 + 64 B) -- H never leaves the SM.
"""
import numpy as np


def flops_per_qp(n, path, as_iters, n_active, admm_iters, nfac, N=16):
    n = np.asarray(n, dtype=np.float64)
    path = np.asarray(path)
    k = np.asarray(n_active, dtype=np.float64)
    build = 0.4e6 * (n / 192.0) ** 2
    f = build + n ** 3 / 3 + 2 * n ** 2
    f = f + 2 * (2 * 144 * N + 24 * n) + 4 * n
    needs_inv = path != 0
    f = f + needs_inv * (n ** 3 / 3)
    f = f + np.asarray(as_iters, dtype=np.float64) * (k ** 2 * n + k ** 3 / 3 + 2 * n ** 2)
    m = 5.0 * n / 3.0
    f = f + np.asarray(admm_iters, dtype=np.float64) * (2 * n ** 2 + 12 * (n + m))
    f = f + np.asarray(nfac, dtype=np.float64) * (build + 2 * n ** 3 / 3)
    return f


def riccati_flops_per_qp(n, N=16):
    """Work of the Riccati pre-pass (csrc/cmpc_riccati*.cuh) for a robot with n stance variables over N stages,
    taking m = n/N variables per stage (convexity makes this a slight under-count):
    per stage  P B 288 m, B'PB 24 m^2, Cholesky + inverse 2 m^3/3, Y = W S and K = W'Y 24 m^2, A'PA 576,
    S'K 288 m, vectors 72 m;  forward sweep 48 n + 24 N;  epilogue as the condensed route."""
    n = np.asarray(n, dtype=np.float64)
    m = n / N
    stage = 2 * m ** 3 / 3 + 48 * m ** 2 + 648 * m + 576
    return N * stage + 48 * n + 24 * N + 2 * (2 * 144 * N + 24 * n) + 4 * n


def wrench_flops_per_qp(sweeps, N=16):
    """Work of the wrench-space projected-Riccati active-set route (csrc/cmpc_wrench.cuh) for a robot that needed
    ``sweeps`` backward + forward sweeps, counted once per robot (the kernel's four threads repeat the 6 x 6
    factorizations; the repeats are not algorithmic work).  Per backward stage, in fused multiply-adds:
    S = P Bbar 180, Gbar = Bbar'S 52, Lam_k / what_k from the four feet 264, ghat and Bbar'q 96, two 6 x 6 Cholesky
    factorizations 112, Nn = I + L'Lam L 217, the five triangular solves of the twelve rows of S and of Bbar'q
    13 * 105 = 1365, the symmetric update P - S Phi S' 468, D A and A'(D A) 150, the vector recursion 96: 3 000 FMA
    = 6.0 kFLOP.  Per forward stage (wrench co-state, four foot projections, multipliers, roll-out) 400 FMA = 0.8 kFLOP.
    The certificate pass (co-states, stationarity, feasibility from first principles) 350 FMA = 0.7 kFLOP per stage."""
    sweeps = np.asarray(sweeps, dtype=np.float64)
    return sweeps * N * (6.0e3 + 0.8e3) + N * 0.7e3 + 1.0e3


def batch_flops(stats, iters, N=16, route_actual=False, prepass=3):
    """Total algorithmic flops of a batch from the (B, NSTAT) stats array and the ADMM iteration counts.
    Robots finished by the Riccati pre-pass (path 4) count at the condensed route's figure for an unconstrained
    robot -- the per-unit figure of SURVEY.md section 8(d), comparable across kernel versions -- unless
    ``route_actual`` asks for the flops of the route they really took."""
    stats = np.asarray(stats)
    path = stats[:, 7].astype(np.int64)
    ric = (path == 4) | (path == 5)
    path = np.where(ric, 0, path)
    admm = (path >= 2)
    f = flops_per_qp(stats[:, 3], path, stats[:, 6], stats[:, 4], np.asarray(iters) * admm, admm.astype(np.float64), N)
    if route_actual:
        f = np.where(ric, riccati_flops_per_qp(stats[:, 3], N), f)
        if prepass == 4:
            p0 = stats[:, 7].astype(np.int64)
            wr = (p0 == 4) | (p0 == 5)
            f = np.where(wr, wrench_flops_per_qp(stats[:, 6] + 1.0, N), f)
    return float(f.sum())


def split_flops(stats, iters, N=16, prepass=4):
    """(flops finished by the pre-pass kernel, flops of the condensed kernel) at the route each robot took."""
    stats = np.asarray(stats)
    p0 = stats[:, 7].astype(np.int64)
    pre = (p0 == 4) | (p0 == 5)
    a = batch_flops(stats[pre], np.asarray(iters)[pre], N, route_actual=True, prepass=prepass) if pre.any() else 0.0
    b = batch_flops(stats[~pre], np.asarray(iters)[~pre], N, route_actual=True, prepass=prepass) if (~pre).any() else 0.0
    return a, b


def bytes_per_qp(N=16):
    return 8 * (12 + 12 * N + 12 * N + 9 + 1 + 1) + 8 + 8 * (12 * N + 28 * N + 12 * N + 12 * N) + 8 + 64
