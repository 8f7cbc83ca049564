"""A NumPy-backed stand-in for the handful of CasADi calls ``convex_mpc/centroidal_mpc.py`` makes, so that the
reference's QP construction (``__init__``, ``_update_sparse_matrix``, ``_compute_bounds``, ``solve_QP``) can be
EXECUTED verbatim in the authoring container, where the real ``casadi`` wheel is not installable.

TEST INFRASTRUCTURE ONLY: used by ``tests/golden/make_golden_qp.py`` to freeze the reference's own QP data
(``H, g, A, lba, uba, lbx, ubx``) into ``tests/golden/reference_qp_vectors.npz``.  It restates nothing of the
reference -- only the container type the reference stores its numbers in:

* ``DM``       a dense float64 matrix with a structural-non-zero pattern;
* ``@``        accumulates over the inner index in ascending order with separately rounded multiply and add,
               the order of CasADi's ``casadi_mtimes`` (column-by-column axpy); for the reference's operands
               every product has at most twelve terms and the zero terms are exact, so the results are the
               values CasADi produces [recall: runtime/casadi_mtimes.hpp];
* ``SX``/``Function``  only as far as ``_create_dynamics_function`` needs: negation, row slices, ``diagcat``;
* ``conic``    records its arguments and returns whatever ``solve_hook`` returns (no solver here).
"""
import numpy as np

inf = np.inf


class Sparsity:
    def __init__(self, nrow, ncol, colind=None, row=None, pattern=None):
        if pattern is not None:
            self.pattern = np.asarray(pattern, dtype=bool)
        else:
            self.pattern = np.zeros((nrow, ncol), dtype=bool)
            colind = np.asarray(colind)
            row = np.asarray(row)
            for c in range(ncol):
                self.pattern[row[colind[c]:colind[c + 1]], c] = True

    def nnz(self):
        return int(self.pattern.sum())

    def size(self):
        return self.pattern.shape


def _as2d(x):
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(-1, 1)          # CasADi turns 1-D input into a column
    return a


class DM:
    def __init__(self, x=None, data=None):
        if isinstance(x, DM):
            self.a, self.nz = x.a.copy(), x.nz.copy()
        elif isinstance(x, Sparsity):
            self.nz = x.pattern.copy()
            self.a = np.zeros(self.nz.shape)
            d = np.asarray(data.a if isinstance(data, DM) else data, dtype=np.float64).reshape(-1)
            # CSC order: column by column, rows ascending
            idx = np.argwhere(self.nz.T)
            for (c, r), v in zip(idx, d):
                self.a[r, c] = v
        else:
            self.a = _as2d(x).copy()
            self.nz = np.ones(self.a.shape, dtype=bool)      # DM(ndarray) is dense

    # ---- constructors
    @staticmethod
    def _mk(a, nz):
        o = DM.__new__(DM)
        o.a, o.nz = a, nz
        return o

    @staticmethod
    def eye(n):
        return DM._mk(np.eye(n), np.eye(n, dtype=bool))

    @staticmethod
    def zeros(r, c=1):
        return DM._mk(np.zeros((r, c)), np.ones((r, c), dtype=bool))

    @staticmethod
    def ones(r, c=1):
        return DM._mk(np.ones((r, c)), np.ones((r, c), dtype=bool))

    @staticmethod
    def triplet(rows, cols, vals, nrow, ncol):
        a = np.zeros((nrow, ncol))
        nz = np.zeros((nrow, ncol), dtype=bool)
        v = np.asarray(vals.a if isinstance(vals, DM) else vals, dtype=np.float64).reshape(-1)
        for r, c, x in zip(rows, cols, v):
            a[r, c] += x
            nz[r, c] = True
        return DM._mk(a, nz)

    # ---- queries
    def size(self):
        return self.a.shape

    @property
    def shape(self):
        return self.a.shape

    def sparsity(self):
        return Sparsity(0, 0, pattern=self.nz)

    def nnz(self):
        return int(self.nz.sum())

    def full(self):
        return self.a.copy()

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    # ---- arithmetic
    def __matmul__(self, other):
        o = other if isinstance(other, DM) else DM(other)
        z = np.zeros((self.a.shape[0], o.a.shape[1]))
        nz = np.zeros(z.shape, dtype=bool)
        for r in range(self.a.shape[1]):                     # ascending inner index, multiply then add
            cols = self.nz[:, r:r + 1] & o.nz[r:r + 1, :]
            if not cols.any():
                continue
            with np.errstate(invalid="ignore"):
                term = self.a[:, r:r + 1] * o.a[r:r + 1, :]
            z = np.where(cols, z + term, z)
            nz |= cols
        return DM._mk(z, nz)

    def __rmatmul__(self, other):
        return DM(other) @ self

    def __add__(self, other):
        o = other if isinstance(other, DM) else DM(np.broadcast_to(_as2d(other), self.a.shape))
        return DM._mk(self.a + o.a, self.nz | o.nz)

    __radd__ = __add__

    def __neg__(self):
        return DM._mk(-self.a, self.nz.copy())

    def __mul__(self, other):
        if isinstance(other, DM):
            return DM._mk(self.a * other.a, self.nz & other.nz)
        return DM._mk(self.a * float(other), self.nz.copy())

    __rmul__ = __mul__

    def __getitem__(self, key):
        return DM._mk(np.atleast_2d(self.a[key]), np.atleast_2d(self.nz[key]))


def vertcat(*xs):
    xs = [x if isinstance(x, DM) else DM(x) for x in xs]
    return DM._mk(np.vstack([x.a for x in xs]), np.vstack([x.nz for x in xs]))


def horzcat(*xs):
    xs = [x if isinstance(x, DM) else DM(x) for x in xs]
    return DM._mk(np.hstack([x.a for x in xs]), np.hstack([x.nz for x in xs]))


def vec(x):
    return DM._mk(x.a.reshape(-1, 1, order="F"), x.nz.reshape(-1, 1, order="F"))


def repmat(x, n, m=1):
    x = x if isinstance(x, DM) else DM(x)
    return DM._mk(np.tile(x.a, (n, m)), np.tile(x.nz, (n, m)))


# ---- the symbolic sliver _create_dynamics_function uses: -sym, sym[a:b, :], diagcat, Function
class SX:
    def __init__(self, fn, shape):
        self.fn, self.shape = fn, shape

    @staticmethod
    def sym(name, r, c=1):
        return SX(lambda env, _n=name: env[_n], (r, c))

    def __neg__(self):
        return SX(lambda env, f=self.fn: -f(env), self.shape)

    def __getitem__(self, key):
        probe = np.zeros(self.shape)[key]
        return SX(lambda env, f=self.fn, k=key: f(env)[k], np.atleast_2d(probe).shape)


def diagcat(*xs):
    def fn(env):
        parts = [x.fn(env) for x in xs]
        R = sum(p.a.shape[0] for p in parts)
        C = sum(p.a.shape[1] for p in parts)
        a = np.zeros((R, C))
        nz = np.zeros((R, C), dtype=bool)
        r = c = 0
        for p in parts:
            h, w = p.a.shape
            a[r:r + h, c:c + w] = p.a
            nz[r:r + h, c:c + w] = p.nz
            r += h
            c += w
        return DM._mk(a, nz)
    return SX(fn, (sum(x.shape[0] for x in xs), sum(x.shape[1] for x in xs)))


class Function:
    def __init__(self, name, ins, outs):
        self.ins, self.outs = ins, outs
        # recover the names the inputs were created with
        self.names = []
        for s in ins:
            probe = {}

            class _Spy(dict):
                def __getitem__(self, k, _p=probe):
                    _p["name"] = k
                    raise KeyError(k)
            try:
                s.fn(_Spy())
            except KeyError:
                pass
            self.names.append(probe["name"])

    def __call__(self, *args):
        env = {n: (a if isinstance(a, DM) else DM(a)) for n, a in zip(self.names, args)}
        return tuple(o.fn(env) for o in self.outs)


# ---- conic: records what the reference hands to the solver
solve_hook = None       # callable(qp_struct, opts, kwargs) -> dict(x=, lam_x=, lam_a=), set by the caller
last_call = {}


class _Conic:
    def __init__(self, name, solver, qp, opts):
        self.name, self.solver, self.qp, self.opts = name, solver, qp, opts

    def __call__(self, **kw):
        last_call.clear()
        last_call.update(kw)
        last_call["_opts"] = self.opts
        last_call["_solver"] = self.solver
        if solve_hook is None:
            n = kw["h"].size()[0]
            m = kw["a"].size()[0]
            return {"x": DM.zeros(n, 1), "lam_x": DM.zeros(n, 1), "lam_a": DM.zeros(m, 1)}
        return solve_hook(self.qp, self.opts, kw)

    def stats(self):
        return {"return_status": "stub"}


def conic(name, solver, qp, opts=None):
    return _Conic(name, solver, qp, opts or {})
