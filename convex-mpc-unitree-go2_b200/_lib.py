"""ctypes binding of ``include/cmpc.h``.  The library is the product: if ``libcmpc.so`` is missing or
cannot be loaded this module raises -- there is no CPU fallback of any kind."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CMPC_LIB") or os.path.join(HERE, "libcmpc.so")   # CMPC_LIB: tools load the timing build

c_int = ctypes.c_int
c_double = ctypes.c_double
c_void_p = ctypes.c_void_p
c_dp = ctypes.POINTER(ctypes.c_double)

NSTAT = 8
STAT_NAMES = ("r_prim", "r_dual", "obj", "n_free", "n_active", "rho", "as_iters", "path")
SOLVED, SOLVED_INACCURATE, MAX_ITER_REACHED, NON_CVX, TOO_MANY_FEET = 1, 2, -2, -7, -20
MODE_ADMM, MODE_ACTIVE_SET = 0, 1

# name -> (restype, argtypes); every symbol declared in include/cmpc.h is listed here
PROTOTYPES = {
    "cmpc_create": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "cmpc_destroy": (c_int, [c_void_p]),
    "cmpc_set_params": (c_int, [c_void_p, c_dp, c_dp, c_double, c_double, c_double, c_double, c_int,
                                c_double, c_double, c_double, c_int, c_int, c_int, c_int]),
    "cmpc_set_max_stance": (c_int, [c_void_p, c_int]),
    "cmpc_set_generic": (c_int, [c_void_p, c_int]),
    "cmpc_leg_jacobian": (c_int, [c_int, c_int, c_void_p, c_void_p, c_dp, c_void_p, c_void_p, c_void_p]),
    "cmpc_set_profile": (c_int, [c_void_p, c_int]),
    "cmpc_last_kernel_ms": (c_int, [c_void_p, c_dp, c_dp]),
    "cmpc_last_kernel_ms3": (c_int, [c_void_p, c_dp, c_dp, c_dp]),
    "cmpc_set_prepass": (c_int, [c_void_p, c_int]),
    "cmpc_workspace_bytes": (c_int, [c_void_p, c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "cmpc_reserve": (c_int, [c_void_p, c_int]),
    "cmpc_contact_table": (c_int, [c_void_p, c_int, c_void_p, c_double, c_double, c_double, c_dp,
                                   c_void_p, c_void_p]),
    "cmpc_generate_traj": (c_int, [c_int, c_int, c_int] + [c_void_p] * 5 + [c_double, c_double, c_double, c_dp, c_dp] +
                           [c_void_p] * 4 + [c_void_p]),
    "cmpc_srb_step": (c_int, [c_int, c_int, c_int] + [c_void_p] * 6 + [c_double, c_dp, c_dp] + [c_void_p] * 4 + [c_void_p]),
    "cmpc_stance_torque": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_double, c_double, c_dp, c_double,
                           c_void_p, c_void_p, c_void_p]),
    "cmpc_pack_contact": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "cmpc_dynamics": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                              c_void_p, c_void_p, c_void_p, c_void_p]),
    "cmpc_build": (c_int, [c_void_p, c_int] + [c_void_p] * 8 + [c_double, c_void_p, c_void_p, c_void_p]),
    "cmpc_solve": (c_int, [c_void_p, c_int] + [c_void_p] * 8 + [c_double, c_void_p, c_int] +
                   [c_void_p] * 8 + [c_void_p]),
    "cmpc_solve_host": (c_int, [c_void_p, c_int] + [c_void_p] * 6 + [c_double, c_double, c_double, c_dp,
                                                                    c_int, c_void_p, c_void_p, c_void_p]),
    "cmpc_cycle_host": (c_int, [c_void_p, c_int] + [c_void_p] * 8 + [c_double, c_double, c_double, c_dp, c_dp,
                                                                    c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cmpc_host_stats": (c_int, [c_void_p, c_int, c_void_p]),
    "cmpc_launch_count": (ctypes.c_longlong, []),
    "cmpc_microbench": (c_int, [c_int, c_dp, c_dp]),
    "cmpc_microbench_dmma": (c_int, [c_int, c_dp, c_dp]),
    "cmpc_microbench_latency": (c_int, [c_int, c_dp]),
    "cmpc_last_error": (ctypes.c_char_p, []),
    "cmpc_version": (ctypes.c_char_p, []),
}

_lib = None


class CmpcError(RuntimeError):
    pass


def load():
    """Load ``libcmpc.so`` (once) and attach prototypes.  Raises if it is absent: build it with
    ``python -c 'import __graft_entry__ as g; g.build()'`` (needs nvcc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CmpcError(
            f"{LIB_PATH} not found: the CUDA library is required (no CPU fallback). "
            "Build it with __graft_entry__.build().")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CmpcError(load().cmpc_last_error().decode())


def darr(values):
    arr = (ctypes.c_double * len(values))(*[float(v) for v in values])
    return arr
