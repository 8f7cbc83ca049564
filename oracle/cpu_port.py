"""ctypes wrapper of ``oracle/libosqp_port.so`` (``osqp_port.c``): the C restatement of the reference's
per-cycle CPU path (contact table -> dynamics -> sparse QP -> OSQP), used as the CPU baseline of
``bench.py`` and as a second checker in ``tests/``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libosqp_port.so")
SRC = os.path.join(HERE, "osqp_port.c")

_lib = None


class PortOpts(ctypes.Structure):
    _fields_ = [("Q", ctypes.c_double * 12), ("R", ctypes.c_double * 12), ("mu", ctypes.c_double),
                ("fz_min", ctypes.c_double), ("eps_abs", ctypes.c_double), ("eps_rel", ctypes.c_double),
                ("rho", ctypes.c_double), ("sigma", ctypes.c_double), ("alpha", ctypes.c_double),
                ("adaptive_rho_tolerance", ctypes.c_double), ("max_iter", ctypes.c_int),
                ("check_termination", ctypes.c_int), ("adaptive_rho_interval", ctypes.c_int),
                ("scaling", ctypes.c_int)]


def build(force=False):
    stale = force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC)
    if stale:
        subprocess.run(["make", "-s", "-C", HERE] + (["-B"] if force else []), check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB)
        assert lib.port_opts_size() == ctypes.sizeof(PortOpts)
        lib.port_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def default_opts(**kw):
    o = PortOpts()
    load().port_default_opts(ctypes.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def solve_batch(rec, opts=None, warm=False, state=None, nthreads=1):
    """Run the port over a ``records.Records`` batch.  Returns dict(w (B,24N), y (B,52N), rho, iters,
    status, obj, nfac).  ``state`` = (w, y, rho) of a previous call for warm starts."""
    lib = load()
    opts = opts or default_opts()
    B, N = rec.B, rec.N
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x0, xr, rf, I, m, t0 = c(rec.x0), c(rec.x_ref), c(rec.r_foot), c(rec.I_world), c(rec.mass), c(rec.t0)
    if state is None:
        w = np.zeros((B, 24 * N)); y = np.zeros((B, 52 * N)); rho = np.zeros(B)
    else:
        w, y, rho = state
    it = np.zeros(B, np.int32); st = np.zeros(B, np.int32); obj = np.zeros(B); nf = np.zeros(B, np.int32)
    off = c([0.5, 0.0, 0.0, 0.5])
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.port_solve_batch(B, N, p(x0), p(xr), p(rf), p(I), p(m), p(t0), ctypes.c_double(rec.dt),
                         ctypes.c_double(rec.gait_hz), ctypes.c_double(rec.duty), p(off), ctypes.byref(opts),
                         int(bool(warm)), int(nthreads), p(w), p(y), p(rho), p(it), p(st), p(obj), p(nf))
    return dict(w=w, y=y, rho=rho, iters=it, status=st, obj=obj, nfac=nf)
