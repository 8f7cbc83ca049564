"""Sources of the ``convex_mpc_b200`` package (see ``convex_mpc_b200/__init__.py`` for the import shim).

  centroidal_mpc.py   batched drop-in for the reference's CentroidalMPC (host side, ctypes)
  records.py          synthetic Go2 MPC records (workload generator)
  _lib.py             ctypes binding of include/cmpc.h
  build.py            nvcc recipe for libcmpc.so (sm_100a)
  csrc/               CUDA kernels and the C-ABI
"""
