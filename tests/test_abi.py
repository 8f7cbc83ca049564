"""The C-ABI library loads on a machine without a GPU and exports every symbol include/cmpc.h declares;
compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from convex_mpc_b200 import _lib, build
    build.build()
    return _lib.load()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cmpc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cmpc_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    from convex_mpc_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cmpc.h but not exported"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == syms
    assert b"sm_100a" in lib.cmpc_version()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.cmpc_create(16, 8, 0, ctypes.byref(h))
    assert rc != 0 and lib.cmpc_last_error() != b""
    from convex_mpc_b200 import _lib
    from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC
    from convex_mpc_b200 import records
    rec = records.random_records(2, seed=1)
    with pytest.raises(_lib.CmpcError):
        CentroidalMPC(None, BatchedComTraj.from_records(rec), verbose=False)


def test_argument_validation(lib):
    h = ctypes.c_void_p()
    assert lib.cmpc_create(0, 8, 0, ctypes.byref(h)) != 0
    assert b"horizon" in lib.cmpc_last_error()
    assert lib.cmpc_create(16, 0, 0, ctypes.byref(h)) != 0
    assert lib.cmpc_set_params(None, None, None, 0.8, 10, 1e-4, 1e-4, 10, 1e-4, 1e-6, 1.6, 1, 0, 10, 25) != 0


def test_module_constants_match_reference():
    import numpy as np
    from convex_mpc_b200 import centroidal_mpc as m
    assert np.array_equal(np.diag(m.COST_MATRIX_Q), [1, 1, 50, 10, 20, 1, 2, 2, 1, 1, 1, 1])
    assert np.array_equal(np.diag(m.COST_MATRIX_R), [1e-5] * 12)
    assert (m.MU, m.NX, m.NU, m.SOLVER_NAME) == (0.8, 12, 12, "osqp")
    o = m.OPTS["osqp"]
    assert (o["eps_abs"], o["eps_rel"], o["max_iter"], o["polish"], o["check_termination"],
            o["adaptive_rho_interval"], o["scaling"]) == (1e-4, 1e-4, 1000, False, 10, 25, 5)
    assert m.OPTS["warm_start_primal"] and m.OPTS["warm_start_dual"]
