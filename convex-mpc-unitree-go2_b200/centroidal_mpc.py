"""Batched, B200-native drop-in for ``convex_mpc/centroidal_mpc.py`` of the reference.

Same constructor, same per-cycle ``solve_QP(go2, traj, verbose)`` call, same module constants and
the same ``sol["x"]`` layout ``w = [x_1..x_N ; u_0..u_{N-1}]`` (centroidal_mpc.py:44,
test_MPC.py:189-192) -- but every ``traj`` field may carry a leading batch dimension of independent
robots, and the QP is built and solved on the GPU by ``libcmpc.so`` (csrc/cmpc.cu) through the C-ABI
of ``include/cmpc.h``.  PyTorch tensors are used as device buffers only.

There is no CPU fallback: importing this module without ``libcmpc.so`` raises.
"""
import ctypes
import time
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import check

# --------------------------------------------------------------------------------
# Model Predictive Control Setting  (same names and values as centroidal_mpc.py:12-38)
# --------------------------------------------------------------------------------
COST_MATRIX_Q = np.diag([1, 1, 50, 10, 20, 1, 2, 2, 1, 1, 1, 1])
COST_MATRIX_R = np.diag([1e-5] * 12)

MU = 0.8
NX = 12
NU = 12
FZ_MIN = 10      # centroidal_mpc.py:127 (hard-coded inside _compute_bounds in the reference)

OPTS = {
    'warm_start_primal': True,
    'warm_start_dual': True,

    "osqp": {
        "eps_abs": 1e-4,
        "eps_rel": 1e-4,
        "max_iter": 1000,
        "polish": False,
        "verbose": False,
        'adaptive_rho': True,
        "check_termination": 10,
        'adaptive_rho_interval': 25,
        "scaling": 5,
        "scaled_termination": True
    }
}

SOLVER_NAME: str = "osqp"        # the reference's CasADi plugin name, kept for importers
BACKEND: str = "cmpc-b200"       # what actually runs here

PHASE_OFFSET = (0.5, 0.0, 0.0, 0.5)   # gait.py:8

_LEG_FIELDS = ("r_fl_foot_world", "r_fr_foot_world", "r_rl_foot_world", "r_rr_foot_world")


class _DM:
    """Tiny stand-in for ``casadi.DM`` so that ``sol["x"].full().flatten()`` (test_MPC.py:190) works."""

    def __init__(self, tensor, batched):
        # a callable is evaluated on first use: assembling the reference's stacked vectors for 65 536 robots moves ~1 GB
        # through HBM per cycle, which a caller that only reads ``sol["u"]`` should not pay for
        self._make = tensor if callable(tensor) else None
        self._tensor = None if callable(tensor) else tensor
        self._batched = batched

    @property
    def tensor(self):
        if self._tensor is None:
            self._tensor = self._make()
            self._make = None
        return self._tensor

    def full(self):
        a = self.tensor.detach().cpu().numpy()
        return a if self._batched else a.reshape(-1, 1)

    def __array__(self, dtype=None):
        a = self.full()
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        return tuple(self.tensor.shape) if self._batched else (self.tensor.numel(), 1)


class MPCSolution(dict):
    """Result of one ``solve_QP``.  Keys of the CasADi dict the reference uses (``x``, ``lam_x``,
    ``lam_a``, ``cost``) plus batched extras (``u`` (B,12,N), ``X`` (B,12,N), ``status``, ``iters``,
    ``r_prim``, ``r_dual``, ``stats``).  Tensors stay on the device until asked for."""


class BatchedComTraj:
    """Duck-typed stand-in for the reference's ``ComTraj`` with a leading batch dimension.

    Carries exactly the members ``solve_QP`` reads under the reference's names:
    ``N, initial_x_vec, compute_x_ref_vec(), contact_table`` and either ``Ad, Bd, gd`` or the raw
    fields ``m, I_com_world, r_*_foot_world`` from which the GPU computes them
    (com_trajectory.py:221-286).  ``time_now`` + ``gait_hz``/``gait_duty`` let the GPU compute the
    contact table as well (gait.py:26-37).
    """

    def __init__(self, N, initial_x_vec, x_ref, dt, *, Ad=None, Bd=None, gd=None, m=None, I_com_world=None,
                 r_foot=None, contact_table=None, time_now=None, gait_hz=3.0, gait_duty=0.6,
                 phase_offset=PHASE_OFFSET):
        self.N = int(N)
        self.initial_x_vec = initial_x_vec
        self._x_ref = x_ref
        self.dt = float(dt)
        self.Ad, self.Bd, self.gd = Ad, Bd, gd
        self.m, self.I_com_world = m, I_com_world
        self.r_foot = r_foot
        if r_foot is not None:
            for i, name in enumerate(_LEG_FIELDS):
                setattr(self, name, r_foot[..., i, :, :])
        self.contact_table = contact_table
        self.time_now = time_now
        self.gait_hz, self.gait_duty, self.phase_offset = gait_hz, gait_duty, tuple(phase_offset)

    def compute_x_ref_vec(self):
        return self._x_ref

    @classmethod
    def from_records(cls, rec, device=None, with_contact_table=False):
        """Wrap a ``records.Records`` batch; ``device`` moves the arrays to that CUDA device."""
        def mv(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            return t.to(device) if device is not None else t
        ct = None
        if with_contact_table:
            from .records import host_contact_table
            ct = mv(host_contact_table(rec.t0, rec.dt, rec.N, rec.gait_hz, rec.duty))
        return cls(rec.N, mv(rec.x0), mv(rec.x_ref), rec.dt, m=mv(rec.mass), I_com_world=mv(rec.I_world),
                   r_foot=mv(rec.r_foot), contact_table=ct, time_now=mv(rec.t0), gait_hz=rec.gait_hz,
                   gait_duty=rec.duty)


class CentroidalMPC:
    """Batched convex MPC.  ``CentroidalMPC(go2, traj)`` then ``solve_QP(go2, traj)`` per cycle.

    ``go2`` is unused, exactly as in the reference (centroidal_mpc.py:41,69).

    Extra keyword arguments (all optional, defaults follow the reference ``OPTS``):
      device      CUDA device (default: current)
      mode        "active_set" (default): exact active-set solve with ADMM fallback;
                  "admm": OSQP-equivalent ADMM only, terminated by eps_abs/eps_rel
      dynamics    "auto" (default): form A_d, B_d, g_d on the GPU from ``traj.m, I_com_world,
                  r_*_foot_world`` when those are present (fast kernel), else use ``traj.Ad/Bd/gd``
                  (generic kernel); "traj" / "device" force one
      eps_abs, eps_rel, max_iter, polish, check_termination, adaptive_rho_interval, rho0, sigma, alpha
      max_stance  upper bound on stance foot-steps per robot (see cmpc_set_max_stance)
      prepass     pre-pass ahead of the condensed kernel: 0 off, 1/2/3 = Riccati sweeps for nominal robots, 4 (default) =
                  wrench-space projected Riccati + primal-dual active set for all robots (cmpc_set_prepass)
      warm_shift  warm start from the previous working set shifted by one horizon stage (the cycle advances by
                  about one dt) instead of the unshifted previous solution the reference re-uses (centroidal_mpc.py:108-110)
      generic_kernel  diagnostics: solve raw-input batches with the generic kernel (the one that serves
                  caller-supplied Ad/Bd/gd) instead of the closed-form fast kernel
    """

    def __init__(self, go2, traj, *, device=None, mode="active_set", dynamics="auto", max_batch=None,
                 eps_abs=None, eps_rel=None, max_iter=None, polish=None, check_termination=None,
                 adaptive_rho_interval=None, rho0=1e-4, sigma=1e-6, alpha=1.6, mu=MU, fz_min=FZ_MIN,
                 Q=None, R=None, max_stance=None, generic_kernel=False, prepass=4, warm_shift=False, verbose=True):
        if not torch.cuda.is_available():
            raise _lib.CmpcError("CentroidalMPC needs a CUDA device (no CPU fallback)")
        self._lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.Q = COST_MATRIX_Q if Q is None else np.asarray(Q, dtype=float)
        self.R = COST_MATRIX_R if R is None else np.asarray(R, dtype=float)
        self.N = int(traj.N)
        self.nvars = self.N * NX + self.N * NU          # centroidal_mpc.py:44
        self.solve_time = 0.0
        self.update_time = 0.0
        self.mode = mode
        self.dynamics = dynamics
        o = OPTS["osqp"]
        self.opts = dict(
            eps_abs=o["eps_abs"] if eps_abs is None else eps_abs,
            eps_rel=o["eps_rel"] if eps_rel is None else eps_rel,
            max_iter=o["max_iter"] if max_iter is None else max_iter,
            polish=o["polish"] if polish is None else polish,
            check_termination=o["check_termination"] if check_termination is None else check_termination,
            adaptive_rho_interval=(o["adaptive_rho_interval"] if o["adaptive_rho"] else 0)
            if adaptive_rho_interval is None else adaptive_rho_interval,
            rho0=rho0, sigma=sigma, alpha=alpha, mu=mu, fz_min=fz_min)
        B = self._batch_of(traj)
        self.max_batch = int(max_batch if max_batch is not None else max(B, 1))
        h = ctypes.c_void_p()
        check(self._lib.cmpc_create(self.N, self.max_batch, self.device.index or 0, ctypes.byref(h)))
        self._h = h
        self._push_params()
        self._auto_stance = None
        if max_stance is None:
            max_stance = self._stance_bound(traj)
            self._auto_stance = max_stance
        if max_stance is not None:
            check(self._lib.cmpc_set_max_stance(self._h, int(max_stance)))
        if generic_kernel:
            check(self._lib.cmpc_set_generic(self._h, 1))
        check(self._lib.cmpc_set_prepass(self._h, int(prepass)))
        self._warm_code = 2 if warm_shift else 1
        self._state_B = None
        self._warm = False
        self._warm_host = 0
        self.kernel_ms = 0.0
        with torch.cuda.device(self.device):
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._prev_shape = None          # x_prev / lam_x_prev / lam_a_prev are None until the first solve (centroidal_mpc.py:62-64)
        self._pending = []
        self.last_stats = None
        if verbose:
            n, N = self.nvars, self.N
            print("\n[QP Init] ===== MPC QP Structure =====")
            print(f"  reference form: vars {n} | constr {28 * N} | horizon N = {N}")
            print(f"  condensed form: vars {12 * N} (stance only) | rows 5 per stance foot-step | backend {BACKEND}")
            print("[QP Init] ✓ Initialization complete.\n")

    # ------------------------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.cmpc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _push_params(self):
        o = self.opts
        Q = _lib.darr(np.diag(self.Q) if np.ndim(self.Q) == 2 else self.Q)
        R = _lib.darr(np.diag(self.R) if np.ndim(self.R) == 2 else self.R)
        mode = {"admm": _lib.MODE_ADMM, "active_set": _lib.MODE_ACTIVE_SET}[self.mode]
        check(self._lib.cmpc_set_params(self._h, Q, R, o["mu"], o["fz_min"], o["eps_abs"], o["eps_rel"],
                                        int(o["max_iter"]), o["rho0"], o["sigma"], o["alpha"], mode,
                                        int(bool(o["polish"])), int(o["check_termination"]),
                                        int(o["adaptive_rho_interval"])))

    def _stance_bound(self, traj):
        """Tight bound on the stance foot-steps of one robot when the GPU computes the contact table itself
        (``traj.time_now`` + gait) and the horizon spans exactly one gait period, as the reference sets it up
        (``N = int(gait_period / dt)``, com_trajectory.py:66): N equally spaced samples on the period circle meet a
        stance arc of length ``duty`` in at most floor(duty N) + 1 points, per leg.  It sizes the shared-memory
        workspace (40 instead of 64 foot-steps for the 3 Hz / 0.6 trot at N = 16: two CTAs per SM).  Arbitrary
        caller-supplied tables keep the general bound 4N."""
        if getattr(traj, "contact_table", None) is not None or getattr(traj, "time_now", None) is None:
            return None
        hz, duty, dt = getattr(traj, "gait_hz", None), getattr(traj, "gait_duty", None), getattr(traj, "dt", None)
        if not hz or duty is None or not dt:
            return None
        period = 1.0 / float(hz)
        if abs(self.N * float(dt) - period) > 1e-9 * period or not (0.0 < float(duty) < 1.0):
            return None
        return min(4 * self.N, 4 * (int(np.floor(float(duty) * self.N + 1e-9)) + 1))

    def set_max_stance(self, nfmax):
        self._auto_stance = None
        check(self._lib.cmpc_set_max_stance(self._h, int(nfmax)))

    def reset(self):
        """Forget the warm-start state (the reference has no such call: it warm-starts forever)."""
        self._warm = False
        self._warm_host = 0
        self._settle_pending()
        self._prev_shape = None

    # ------------------------------------------------------------------------------------------
    def _batch_of(self, traj):
        x0 = traj.initial_x_vec
        shp = tuple(x0.shape)
        if len(shp) == 1 or (len(shp) == 2 and shp == (12, 1)):
            return 0       # un-batched (reference shapes)
        return int(shp[0])

    def _dev(self, a, shape):
        """Device FP64 contiguous tensor of ``shape`` from numpy / torch input (no copy if possible)."""
        if isinstance(a, torch.Tensor):
            t = a
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        return t.reshape(shape).contiguous()

    def _alloc_state(self, B):
        if self._state_B == B:
            return
        N, dev = self.N, self.device
        if B > self.max_batch:
            raise _lib.CmpcError(f"batch {B} exceeds max_batch {self.max_batch} given at construction")
        f64 = dict(dtype=torch.float64, device=dev)
        self._u = torch.zeros(B, 12 * N, **f64)
        self._y = torch.zeros(B, 28 * N, **f64)
        self._rho = torch.zeros(B, **f64)
        self._X = torch.empty(B, 12 * N, **f64)
        self._nu = torch.empty(B, 12 * N, **f64)
        self._stats = torch.empty(B, _lib.NSTAT, **f64)
        self._status = torch.empty(B, dtype=torch.int32, device=dev)
        self._iters = torch.empty(B, dtype=torch.int32, device=dev)
        self._mask = torch.empty(B, (4 * N + 63) // 64, dtype=torch.int64, device=dev)
        self._state_B = B
        self._warm = False

    def _gather(self, traj, B, stream):
        """Collect device pointers for one call.  Returns (dict of tensors kept alive, use_AdBd)."""
        N = self.N
        if self._auto_stance is not None and self._stance_bound(traj) != self._auto_stance:
            # the gait the automatic bound was derived from no longer holds: back to the general bound 4N
            self._auto_stance = None
            check(self._lib.cmpc_set_max_stance(self._h, 4 * N))
        t = {}
        t["x0"] = self._dev(traj.initial_x_vec, (B, 12))
        t["x_ref"] = self._dev(traj.compute_x_ref_vec(), (B, 12, N))
        have_ab = all(getattr(traj, k, None) is not None for k in ("Ad", "Bd", "gd"))
        have_raw = (getattr(traj, "m", None) is not None and getattr(traj, "I_com_world", None) is not None and
                    (getattr(traj, "r_foot", None) is not None or
                     all(getattr(traj, name, None) is not None for name in _LEG_FIELDS)))
        # "auto": the raw fields win when present -- the reference's own ComTraj carries both (com_trajectory.py:
        # 39-40,204-207 next to Ad/Bd/gd), and from the raw fields the GPU forms the same A_d, B_d, g_d in closed
        # form and takes the fast kernel; Ad/Bd/gd alone (arbitrary structure) go to the generic kernel
        use_ab = (have_ab and not have_raw) if self.dynamics == "auto" else (self.dynamics == "traj")
        if use_ab:
            if not have_ab:
                raise _lib.CmpcError("dynamics='traj' needs traj.Ad, traj.Bd and traj.gd")
            t["Ad"] = self._dev(traj.Ad, (B, 12, 12))
            t["Bd"] = self._dev(traj.Bd, (B, N, 12, 12))
            t["gd"] = self._dev(traj.gd, (B, 12))
        else:
            if not have_raw:
                raise _lib.CmpcError("dynamics='device' needs traj.m, traj.I_com_world and the foot lever arms")
            m = traj.m
            if not isinstance(m, torch.Tensor) and np.ndim(m) == 0:
                m = np.full(B, float(m))
            t["mass"] = self._dev(m, (B,))
            t["I_world"] = self._dev(traj.I_com_world, (B, 3, 3))
            rf = getattr(traj, "r_foot", None)
            if rf is None:
                legs = [getattr(traj, name) for name in _LEG_FIELDS]
                if isinstance(legs[0], torch.Tensor):
                    rf = torch.stack([l.reshape(B, 3, N) for l in legs], dim=1)
                else:
                    rf = np.stack([np.asarray(l, dtype=np.float64).reshape(B, 3, N) for l in legs], axis=1)
            t["r_foot"] = self._dev(rf, (B, 4, 3, N))
            dt = getattr(traj, "dt", None)
            if dt is None and have_ab:
                # the reference does not store its time step; A_d = I + dt A_c has dt at [0, 6] (com_trajectory.py:234-239,278)
                Ad0 = traj.Ad
                dt = float(Ad0.reshape(-1, 12, 12)[0, 0, 6]) if not isinstance(Ad0, torch.Tensor) else float(Ad0.reshape(-1, 12, 12)[0, 0, 6].item())
            if dt is None:
                dt = (1.0 / getattr(traj, "gait_hz", 3.0)) / N
            t["dt"] = float(dt)
        # contact mask: the table the reference carries (com_trajectory.py:106), else computed here
        ct = getattr(traj, "contact_table", None)
        if ct is not None:
            if isinstance(ct, torch.Tensor):
                c = ct.to(device=self.device, dtype=torch.int32).reshape(B, 4, N).contiguous()
            else:
                c = torch.from_numpy(np.ascontiguousarray(np.asarray(ct).astype(np.int32))).to(self.device)
                c = c.reshape(B, 4, N).contiguous()
            t["ct"] = c
            check(self._lib.cmpc_pack_contact(self._h, B, c.data_ptr(), self._mask.data_ptr(), stream))
        else:
            tn = getattr(traj, "time_now", None)
            if tn is None:
                raise _lib.CmpcError("traj needs contact_table or time_now (+ gait_hz, gait_duty)")
            if not isinstance(tn, torch.Tensor) and np.ndim(tn) == 0:
                tn = np.full(B, float(tn))
            t["t0"] = self._dev(tn, (B,))
            dt = float(getattr(traj, "dt", None) or (1.0 / traj.gait_hz) / N)
            off = _lib.darr(getattr(traj, "phase_offset", PHASE_OFFSET))
            check(self._lib.cmpc_contact_table(self._h, B, t["t0"].data_ptr(), dt, float(traj.gait_hz),
                                               float(traj.gait_duty), off, self._mask.data_ptr(), stream))
        return t, use_ab

    # ------------------------------------------------------------------------------------------
    def _stacked(self, B, N, sq):
        """Makers of the reference's stacked vectors (centroidal_mpc.py:108-110: x (24N), lam_x (24N), lam_a (28N)) from
        the device buffers; evaluated on first use."""
        X, u, y, nu, dev = self._X, self._u, self._y, self._nu, self.device
        return {"x": lambda: sq(torch.cat([X[:B], u[:B]], dim=1)),
                "lam_x": lambda: sq(torch.cat([torch.zeros(B, 12 * N, dtype=torch.float64, device=dev), y[:B, :12 * N]], dim=1)),
                "lam_a": lambda: sq(torch.cat([nu[:B], y[:B, 12 * N:]], dim=1))}

    def _settle_pending(self):
        for ref in getattr(self, "_pending", ()):
            d = ref()
            if d is not None:
                d.tensor
        self._pending = []

    def _prev(self, key):
        shp = getattr(self, "_prev_shape", None)
        if shp is None:
            return None
        B, N, batched = shp
        d = _DM(self._stacked(B, N, (lambda a: a) if batched else (lambda a: a[0]))[key], batched)
        self._pending.append(weakref.ref(d))
        return d

    # the reference keeps the raw previous solution for the warm start (centroidal_mpc.py:108-110); here the device
    # buffers themselves are the warm start, the stacked views are built when somebody asks
    x_prev = property(lambda self: self._prev("x"))
    lam_x_prev = property(lambda self: self._prev("lam_x"))
    lam_a_prev = property(lambda self: self._prev("lam_a"))

    # ------------------------------------------------------------------------------------------
    def solve_QP(self, go2, traj, verbose: bool = False):
        """One MPC cycle for every robot in ``traj`` (centroidal_mpc.py:69-120)."""
        t0 = time.perf_counter()
        B0 = self._batch_of(traj)
        B = max(B0, 1)
        N = self.N
        if int(traj.N) != N:
            raise _lib.CmpcError(f"traj.N = {traj.N} but the solver was built for N = {N}")
        self._alloc_state(B)
        with torch.cuda.device(self.device):
            self._settle_pending()          # stacked vectors of the previous solution somebody still holds are built now
            stream = torch.cuda.current_stream().cuda_stream
            t, use_ab = self._gather(traj, B, stream)
            # (no synchronisation here: the solve is enqueued behind the contact-table kernel on the same stream; update_time
            # is the host time of the update, as the reference's is, centroidal_mpc.py:73-96)
            t1 = time.perf_counter()
            p = lambda k: t[k].data_ptr() if k in t else None
            self._ev[0].record()
            check(self._lib.cmpc_solve(
                self._h, B, p("Ad"), p("Bd"), p("gd"), p("x0"), p("x_ref"), p("r_foot"), p("I_world"), p("mass"),
                float(t.get("dt", 0.0)), self._mask.data_ptr(), self._warm_code if self._warm else 0,
                self._u.data_ptr(), self._y.data_ptr(), self._rho.data_ptr(), self._X.data_ptr(),
                self._nu.data_ptr(), self._status.data_ptr(), self._iters.data_ptr(), self._stats.data_ptr(),
                stream))
            self._ev[1].record()
            torch.cuda.current_stream().synchronize()   # the reference call is blocking (centroidal_mpc.py:98)
            self.kernel_ms = self._ev[0].elapsed_time(self._ev[1])     # device time of the fused solve kernel
        t2 = time.perf_counter()
        self.update_time = (t1 - t0) * 1e3      # ms, as centroidal_mpc.py:102-105
        self.solve_time = (t2 - t1) * 1e3
        if OPTS.get("warm_start_primal", True):
            self._warm = True

        batched = B0 > 0
        sq = (lambda a: a) if batched else (lambda a: a[0])
        sol = MPCSolution()
        mk = self._stacked(B, N, sq)
        sol["x"], sol["lam_x"], sol["lam_a"] = (_DM(mk[k], batched) for k in ("x", "lam_x", "lam_a"))
        self._pending = [weakref.ref(sol[k]) for k in ("x", "lam_x", "lam_a")]
        stats_ = self._stats
        sol["cost"] = _DM(lambda: sq(stats_[:B, 2:3].clone()), batched)
        self._pending.append(weakref.ref(sol["cost"]))
        sol["u"] = sq(self._u.view(B, N, 12).transpose(1, 2))       # (B,12,N): U_opt of test_MPC.py:192
        sol["X"] = sq(self._X.view(B, N, 12).transpose(1, 2))
        sol["status"] = sq(self._status)
        sol["iters"] = sq(self._iters)
        sol["stats"] = sq(self._stats)
        sol["r_prim"] = sq(self._stats[:, 0])
        sol["r_dual"] = sq(self._stats[:, 1])
        self._prev_shape = (B, N, batched)
        self.last_stats = self._stats
        if verbose:
            st = self._status.cpu().numpy()
            print(f"[QP SOLVER] update (gather + contact mask) takes {self.update_time:.3f} ms")
            print(f"[QP SOLVER] solver takes {self.solve_time:.3f} ms for {B} robot(s)")
            tot = (t2 - t0)
            print(f"[QP SOLVER] total time = {tot * 1e3:.3f} ms  ({B / tot:.1f} QPs/s)")
            print(f"[QP SOLVER] status: solved {int((st == 1).sum())}/{B}, "
                  f"max iters {int(self._iters.max())}")
        return sol

    # ------------------------------------------------------------------------------------------
    def enqueue(self, traj, stream=None):
        """Asynchronous variant of ``solve_QP`` for device-resident loops: enqueue the contact table and the fused
        solve of one cycle on ``stream`` (default: torch's current stream) and return at once -- no host
        synchronisation, no events, no solution dict, so the call can be captured in a CUDA graph
        (tools/closed_loop.py).  Results land in ``self._u`` (B, 12N: forces, entries 0..11 = step 0), ``self._X``,
        ``self._y``, ``self._nu``, ``self._status``, ``self._iters``, ``self._stats``; tensors of ``traj`` must stay
        alive until the stream has run."""
        B = max(self._batch_of(traj), 1)
        if int(traj.N) != self.N:
            raise _lib.CmpcError(f"traj.N = {traj.N} but the solver was built for N = {self.N}")
        self._alloc_state(B)
        with torch.cuda.device(self.device):
            s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
            t, _ = self._gather(traj, B, s)
            p = lambda k: t[k].data_ptr() if k in t else None
            check(self._lib.cmpc_solve(
                self._h, B, p("Ad"), p("Bd"), p("gd"), p("x0"), p("x_ref"), p("r_foot"), p("I_world"), p("mass"),
                float(t.get("dt", 0.0)), self._mask.data_ptr(), self._warm_code if self._warm else 0,
                self._u.data_ptr(), self._y.data_ptr(), self._rho.data_ptr(), self._X.data_ptr(),
                self._nu.data_ptr(), self._status.data_ptr(), self._iters.data_ptr(), self._stats.data_ptr(), s))
        self._keep = t
        if OPTS.get("warm_start_primal", True):
            self._warm = True

    # ------------------------------------------------------------------------------------------
    def solve_host(self, x0, x_ref, r_foot, I_world, mass, t0, dt, gait_hz=3.0, duty=0.6,
                   phase_offset=PHASE_OFFSET, out=None):
        """Whole cycle on HOST buffers through ``cmpc_solve_host`` (chunked, copy/compute overlapped).

        Inputs: C-contiguous float64 numpy arrays or CPU torch tensors (pinned memory gives true
        overlap).  Returns ``(u (B,12N), status (B), iters (B))`` as numpy views of ``out`` buffers.
        """
        def host_ptr(a, n):
            if isinstance(a, torch.Tensor):
                assert a.device.type == "cpu" and a.dtype == torch.float64 and a.is_contiguous() and a.numel() == n
                return a.data_ptr()
            a = np.asarray(a)
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == n
            return a.ctypes.data
        N = self.N
        B = int(np.prod(tuple(mass.shape)))
        if out is None:
            out = (torch.empty(B, 12 * N, dtype=torch.float64).pin_memory(),
                   torch.empty(B, dtype=torch.int32).pin_memory(),
                   torch.empty(B, dtype=torch.int32).pin_memory())
        u, st, it = out
        check(self._lib.cmpc_solve_host(
            self._h, B, host_ptr(x0, B * 12), host_ptr(x_ref, B * 12 * N), host_ptr(r_foot, B * 12 * N),
            host_ptr(I_world, B * 9), host_ptr(mass, B), host_ptr(t0, B), float(dt), float(gait_hz), float(duty),
            _lib.darr(phase_offset), int(self._warm_host),
            u.data_ptr(), st.data_ptr(), it.data_ptr()))
        self._warm_host = 1 if OPTS.get("warm_start_primal", True) else 0
        return u, st, it

    def cycle_host(self, x0, R_world_to_body, foot_lever, cmd, t0, pos_des, I_world, mass, dt, hip_offset,
                   gait_hz=3.0, duty=0.6, phase_offset=PHASE_OFFSET, first_step_only=False, out=None):
        """One whole MPC cycle from HOST state + command through ``cmpc_cycle_host``: reference trajectory and lever arms
        (``ComTraj.generate_traj``, com_trajectory.py:27-211), contact table, QP solve, and the slice
        ``U_opt = w[12N:]`` of test_MPC.py:189-196 -- 408 bytes in per robot, 96 bytes out with ``first_step_only``.

        Inputs: C-contiguous float64 numpy arrays or CPU torch tensors (pinned memory gives true overlap);
        ``pos_des`` (B,3) is updated in place.  Returns ``(u, status, iters)`` (``u`` (B,12) or (B,12N))."""
        def host_ptr(a, n):
            if isinstance(a, torch.Tensor):
                assert a.device.type == "cpu" and a.dtype == torch.float64 and a.is_contiguous() and a.numel() == n
                return a.data_ptr()
            a = np.asarray(a)
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == n
            return a.ctypes.data
        N = self.N
        B = int(np.prod(tuple(mass.shape)))
        w = 12 if first_step_only else 12 * N
        if out is None:
            out = (torch.empty(B, w, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory(),
                   torch.empty(B, dtype=torch.int32).pin_memory())
        u, st, it = out
        assert u.numel() == B * w
        check(self._lib.cmpc_cycle_host(
            self._h, B, host_ptr(x0, B * 12), host_ptr(R_world_to_body, B * 9), host_ptr(foot_lever, B * 12),
            host_ptr(cmd, B * 4), host_ptr(t0, B), host_ptr(pos_des, B * 3), host_ptr(I_world, B * 9), host_ptr(mass, B),
            float(dt), float(gait_hz), float(duty), _lib.darr(phase_offset), _lib.darr(np.asarray(hip_offset, dtype=np.float64).reshape(-1)),
            int(self._warm_host), int(bool(first_step_only)), u.data_ptr(), st.data_ptr(), it.data_ptr()))
        self._warm_host = 1 if OPTS.get("warm_start_primal", True) else 0
        return u, st, it

    def host_stats(self, B):
        s = np.empty((B, _lib.NSTAT))
        check(self._lib.cmpc_host_stats(self._h, B, s.ctypes.data))
        return s
