"""Shared test helpers: oracle QP construction from records, and the host emulation of the kernel
source (tests/_emul) used by the CPU suite."""
import ctypes
import os
import subprocess

import numpy as np

from oracle import condensed_qp, dynamics_ref, exact, gait_ref, sparse_qp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_DIR = os.path.join(ROOT, "tests", "_emul")
EMUL_SRC = os.path.join(EMUL_DIR, "cmpc_emul.cpp")
EMUL_LIB = os.path.join(EMUL_DIR, "libcmpc_emul.so")
CORE = os.path.join(ROOT, "convex-mpc-unitree-go2_b200", "csrc", "cmpc_core.cuh")
FAST = os.path.join(ROOT, "convex-mpc-unitree-go2_b200", "csrc", "cmpc_fast.cuh")
RIC = os.path.join(ROOT, "convex-mpc-unitree-go2_b200", "csrc", "cmpc_riccati.cuh")
TRAJ = os.path.join(ROOT, "convex-mpc-unitree-go2_b200", "csrc", "cmpc_traj.cuh")
WRENCH = os.path.join(ROOT, "convex-mpc-unitree-go2_b200", "csrc", "cmpc_wrench.cuh")

PHASE_OFFSET = np.array([0.5, 0.0, 0.0, 0.5])


def oracle_inputs(rec, b):
    """(contact, Ad, Bd, gd) of robot b by the oracle's restatement of the reference."""
    ct = gait_ref.contact_table(rec.t0[b], rec.dt, rec.N, rec.gait_hz, rec.duty)
    yaw = dynamics_ref.yaw_average(rec.x_ref[b])
    Ac, Bc, gc = dynamics_ref.continuous_dynamics(rec.mass[b], rec.I_world[b], yaw, rec.r_foot[b])
    Ad, Bd, gd = dynamics_ref.discrete_dynamics_closed(Ac, Bc, gc, rec.dt)
    return ct, Ad, Bd, gd


def oracle_solution(rec, b, contact=None):
    ct, Ad, Bd, gd = oracle_inputs(rec, b)
    if contact is not None:
        ct = contact
    cq = condensed_qp.build(Ad, Bd, gd, rec.x0[b], rec.x_ref[b], ct)
    sol = exact.solve_exact(cq["H"], cq["g"], cq["A"], cq["l"], cq["u"])
    return dict(ct=ct, Ad=Ad, Bd=Bd, gd=gd, cq=cq, sol=sol)


def force_error(u, u_star):
    """(abs inf-norm error in N, error relative to the 1e-2 N + 1e-3 rel tolerance of the north star)."""
    d = np.abs(u - u_star)
    tol = 1e-2 + 1e-3 * np.abs(u_star)
    return d.max(), (d / tol).max()


# ------------------------------------------------------------------------------------------------
# host emulation of csrc/cmpc_core.cuh (one-thread CTA)
# ------------------------------------------------------------------------------------------------
def build_emul():
    stale = (not os.path.exists(EMUL_LIB) or
             os.path.getmtime(EMUL_LIB) < max(os.path.getmtime(EMUL_SRC), os.path.getmtime(CORE), os.path.getmtime(FAST), os.path.getmtime(RIC), os.path.getmtime(TRAJ), os.path.getmtime(WRENCH)))
    if stale:
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", EMUL_LIB, EMUL_SRC],
                       check=True)
    return ctypes.CDLL(EMUL_LIB)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Emul:
    def __init__(self):
        self.lib = build_emul()
        self.lib.emul_ws_bytes.restype = ctypes.c_size_t
        self.psize = self.lib.emul_params_size()

    def params(self, **kw):
        buf = (ctypes.c_char * self.psize)()
        self.lib.emul_default_params(buf)
        d = np.frombuffer(buf, dtype=np.float64, count=31)
        i = np.frombuffer(buf, dtype=np.int32, offset=31 * 8, count=6)
        names_d = {"mu": 24, "fz_min": 25, "eps_abs": 26, "eps_rel": 27, "rho0": 28, "sigma": 29, "alpha": 30}
        names_i = {"max_iter": 0, "mode": 1, "polish": 2, "check_termination": 3,
                   "adaptive_rho_interval": 4, "pdas_max_iter": 5}
        for k, v in kw.items():
            if k in names_d:
                d[names_d[k]] = v
            else:
                i[names_i[k]] = v
        return buf

    def contact_table(self, t0, dt, N, hz, duty, off=PHASE_OFFSET):
        t0 = np.ascontiguousarray(t0, dtype=np.float64)
        B = t0.shape[0]
        mask = np.zeros((B, (4 * N + 63) // 64), np.uint64)
        off = np.ascontiguousarray(off, dtype=np.float64)
        self.lib.emul_contact_table(B, N, _p(t0), ctypes.c_double(dt), ctypes.c_double(hz), ctypes.c_double(duty),
                                    _p(off), _p(mask))
        return mask

    def build(self, rec, Ad=None, Bd=None, gd=None, **kw):
        B, N = rec.B, rec.N
        n = 12 * N
        H = np.zeros((B, n, n))
        g = np.zeros((B, n))
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        Ad, Bd, gd = c(Ad), c(Bd), c(gd)
        self.lib.emul_build(self.params(**kw), B, N, _p(Ad), _p(Bd), _p(gd), _p(rec.x0), _p(rec.x_ref),
                            _p(rec.r_foot), _p(rec.I_world), _p(rec.mass), ctypes.c_double(rec.dt), _p(H), _p(g))
        return H, g

    def solve(self, rec, mask=None, nfmax=None, warm=0, state=None, Ad=None, Bd=None, gd=None, **kw):
        B, N = rec.B, rec.N
        nfmax = 4 * N if nfmax is None else nfmax
        if mask is None:
            mask = self.contact_table(rec.t0, rec.dt, N, rec.gait_hz, rec.duty)
        if state is None:
            u = np.zeros((B, 12 * N)); y = np.zeros((B, 28 * N)); rho = np.zeros(B)
        else:
            u, y, rho = state
        X = np.zeros((B, 12 * N)); nu = np.zeros((B, 12 * N))
        st = np.zeros(B, np.int32); it = np.zeros(B, np.int32); stats = np.zeros((B, 8))
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        Ad, Bd, gd = c(Ad), c(Bd), c(gd)
        self.lib.emul_solve(self.params(**kw), B, N, nfmax, _p(Ad), _p(Bd), _p(gd), _p(rec.x0), _p(rec.x_ref),
                            _p(rec.r_foot), _p(rec.I_world), _p(rec.mass), ctypes.c_double(rec.dt), _p(mask), warm,
                            _p(u), _p(y), _p(rho), _p(X), _p(nu), _p(st), _p(it), _p(stats))
        return dict(u=u, y=y, rho=rho, X=X, nu=nu, status=st, iters=it, stats=stats, mask=mask)

    def solve_fast(self, rec, mask=None, nfmax=None, warm=0, state=None, **kw):
        """v2 raw-input path (csrc/cmpc_fast.cuh) under the same one-thread-CTA emulation."""
        B, N = rec.B, rec.N
        nfmax = 4 * N if nfmax is None else nfmax
        if mask is None:
            mask = self.contact_table(rec.t0, rec.dt, N, rec.gait_hz, rec.duty)
        if state is None:
            u = np.zeros((B, 12 * N)); y = np.zeros((B, 28 * N)); rho = np.zeros(B)
        else:
            u, y, rho = state
        X = np.zeros((B, 12 * N)); nu = np.zeros((B, 12 * N))
        st = np.zeros(B, np.int32); it = np.zeros(B, np.int32); stats = np.zeros((B, 8))
        self.lib.emul_solve_fast(self.params(**kw), B, N, nfmax, _p(rec.x0), _p(rec.x_ref), _p(rec.r_foot),
                                 _p(rec.I_world), _p(rec.mass), ctypes.c_double(rec.dt), _p(mask), warm,
                                 _p(u), _p(y), _p(rho), _p(X), _p(nu), _p(st), _p(it), _p(stats))
        return dict(u=u, y=y, rho=rho, X=X, nu=nu, status=st, iters=it, stats=stats, mask=mask)

    def riccati(self, rec, mask=None, nfmax=None, warm=0, state=None, **kw):
        """Riccati pre-pass (csrc/cmpc_riccati.cuh); ``done`` marks the robots it finished."""
        B, N = rec.B, rec.N
        nfmax = 4 * N if nfmax is None else nfmax
        if mask is None:
            mask = self.contact_table(rec.t0, rec.dt, N, rec.gait_hz, rec.duty)
        if state is None:
            u = np.zeros((B, 12 * N)); y = np.zeros((B, 28 * N)); rho = np.zeros(B)
        else:
            u, y, rho = state
        X = np.zeros((B, 12 * N)); nu = np.zeros((B, 12 * N))
        st = np.zeros(B, np.int32); it = np.zeros(B, np.int32); stats = np.zeros((B, 8))
        done = np.zeros(B, np.int32)
        self.lib.emul_riccati(self.params(**kw), B, N, nfmax, _p(rec.x0), _p(rec.x_ref), _p(rec.r_foot),
                              _p(rec.I_world), _p(rec.mass), ctypes.c_double(rec.dt), _p(mask), warm,
                              _p(u), _p(y), _p(rho), _p(X), _p(nu), _p(st), _p(it), _p(stats), _p(done))
        return dict(u=u, y=y, rho=rho, X=X, nu=nu, status=st, iters=it, stats=stats, mask=mask, done=done)

    def leg_jacobian(self, q, R_wb, link):
        """traj::leg_jacobian (csrc/cmpc_traj.cuh) on the host: (J (B,4,3,3), foot_pos_body (B,4,3))."""
        q = np.ascontiguousarray(q, dtype=np.float64); R_wb = np.ascontiguousarray(R_wb, dtype=np.float64)
        B = q.shape[0]
        J = np.zeros((B, 4, 3, 3)); pb = np.zeros((B, 4, 3))
        self.lib.emul_leg_jacobian(B, _p(q), _p(R_wb), _p(np.asarray(link, dtype=np.float64)), _p(J), _p(pb))
        return J, pb

    def wrench(self, rec, mask=None, nfmax=None, warm=0, state=None, **kw):
        """Wrench-space projected Riccati + PDAS (csrc/cmpc_wrench.cuh); ``done`` marks the robots it finished,
        ``sweeps`` the Riccati sweeps it ran."""
        B, N = rec.B, rec.N
        nfmax = 4 * N if nfmax is None else nfmax
        if mask is None:
            mask = self.contact_table(rec.t0, rec.dt, N, rec.gait_hz, rec.duty)
        if state is None:
            u = np.zeros((B, 12 * N)); y = np.zeros((B, 28 * N)); rho = np.zeros(B)
        else:
            u, y, rho = state
        X = np.zeros((B, 12 * N)); nu = np.zeros((B, 12 * N))
        st = np.zeros(B, np.int32); it = np.zeros(B, np.int32); stats = np.zeros((B, 8))
        done = np.zeros(B, np.int32); sweeps = np.zeros(B, np.int32)
        self.lib.emul_wrench(self.params(**kw), B, N, nfmax, _p(rec.x0), _p(rec.x_ref), _p(rec.r_foot),
                             _p(rec.I_world), _p(rec.mass), ctypes.c_double(rec.dt), _p(mask), warm,
                             _p(u), _p(y), _p(rho), _p(X), _p(nu), _p(st), _p(it), _p(stats), _p(done), _p(sweeps))
        return dict(u=u, y=y, rho=rho, X=X, nu=nu, status=st, iters=it, stats=stats, mask=mask, done=done, sweeps=sweeps)

    def generate_traj(self, N, x0, R_wb, lever, cmd, t0, dt, hz, duty, hip, pos_des, off=PHASE_OFFSET):
        """csrc/cmpc_traj.cuh under the host emulation: returns (pos_des_out, x_ref (B,12,N), r_foot (B,4,3,N))."""
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x0, R_wb, lever, cmd, t0, hip, pos_des, off = map(c, (x0, R_wb, lever, cmd, t0, hip, pos_des, off))
        B = x0.shape[0]
        out_pd = np.zeros((B, 3)); x_ref = np.zeros((B, 12, N)); r_foot = np.zeros((B, 4, 3, N))
        self.lib.emul_generate_traj(N, B, _p(x0), _p(R_wb), _p(lever), _p(cmd), _p(t0), ctypes.c_double(dt),
                                    ctypes.c_double(hz), ctypes.c_double(duty), _p(off), _p(hip), _p(pos_des),
                                    _p(out_pd), _p(x_ref), _p(r_foot))
        return out_pd, x_ref, r_foot


def golden_traj_batches(path=None):
    """The reference-generated trajectory cases (tests/golden/make_golden_traj.py) grouped by (hz, duty, N)."""
    path = path or os.path.join(ROOT, "tests", "golden", "reference_traj_vectors.npz")
    g = np.load(path)
    groups = {}
    for i in range(int(g["count"])):
        k = f"tr{i}_"
        cfg = tuple(g[k + "cfg"])
        groups.setdefault(cfg, []).append({f: g[k + f] for f in ("x0", "R_wb", "levers", "mass", "inertia", "pos_des_in", "cmd",
                                                                "t_now", "pos_des_out", "x_ref", "contact", "r_foot", "Ad", "Bd", "gd")})
    out = []
    for (hz, duty, N, dt), cases in groups.items():
        st = lambda f: np.stack([c[f] for c in cases])
        out.append(dict(hz=float(hz), duty=float(duty), N=int(N), dt=float(dt), hip=g["hip"],
                        **{f: st(f) for f in cases[0]}))
    return out
