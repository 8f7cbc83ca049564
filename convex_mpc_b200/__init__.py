"""Import name of the package whose sources live in ``convex-mpc-unitree-go2_b200/`` (that directory
name is not a Python identifier).  ``import convex_mpc_b200`` == the B200-native batched drop-in for
the reference's ``convex_mpc/centroidal_mpc.py``."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                              "convex-mpc-unitree-go2_b200"))
__version__ = "0.1.0"

from . import records  # noqa: E402,F401  (NumPy only)


def __getattr__(name):
    # the solver module needs torch + libcmpc.so: import it on first use so that record generation
    # and build tooling work on machines without a GPU
    if name in ("CentroidalMPC", "BatchedComTraj", "MPCSolution", "COST_MATRIX_Q", "COST_MATRIX_R", "MU", "NX", "NU",
                "OPTS", "SOLVER_NAME", "centroidal_mpc"):
        import importlib
        _m = importlib.import_module(__name__ + ".centroidal_mpc")
        return _m if name == "centroidal_mpc" else getattr(_m, name)
    raise AttributeError(name)
