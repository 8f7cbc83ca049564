"""Live split of one cmpc_solve call at the bench's size into its kernels (cmpc_set_profile / cmpc_last_kernel_ms3): per solve
(sweep kernel incl. robot constants, certificate kernel, what is left of the condensed kernel after the join, whole call) in ms.
Development aid: python tools/kernel_split.py"""
import sys, os, ctypes
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from convex_mpc_b200 import records, _lib
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC
lib = _lib.load()
rec = records.random_records(65536, seed=65536)
traj = BatchedComTraj.from_records(rec, device="cuda:0")
mpc = CentroidalMPC(None, traj, verbose=False, max_stance=40)
_lib.check(lib.cmpc_set_profile(mpc._h, 1))
a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
out = []
for i in range(8):
    mpc.reset(); mpc.solve_QP(None, traj)
    lib.cmpc_last_kernel_ms3(mpc._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    out.append((round(a.value, 3), round(b.value, 3), round(c.value, 3), round(mpc.kernel_ms, 3)))
print(out[3:])
