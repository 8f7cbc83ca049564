"""Per-phase cycle accounting of the fast kernel (needs the timing build):
    python convex-mpc-unitree-go2_b200/build.py --timing
    CMPC_LIB=convex-mpc-unitree-go2_b200/libcmpc_timing.so python tools/phase_times.py [B] [stress] [mode]
"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import _lib, records
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC

NAMES = ["setup", "vectors", "build_H", "cholesky", "backsolve+viol", "trtri", "active_set", "admm", "polish", "rollout", "residuals", "outputs", "admm:factor", "admm:iteration", "admm:check", "sub15"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
stress = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
mode = sys.argv[3] if len(sys.argv) > 3 else "active_set"
lib = _lib.load()
rec = records.random_records(B, seed=65536, stress=stress)
traj = BatchedComTraj.from_records(rec, device="cuda:0")
kw = dict(eps_abs=1e-5, eps_rel=1e-5, max_iter=4000) if mode == "admm" else {}
mpc = CentroidalMPC(None, traj, verbose=False, max_stance=40, mode=mode, **kw)
for _ in range(2):
    mpc.reset(); mpc.solve_QP(None, traj)
out = (ctypes.c_double * 32)()
lib.cmpc_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.cmpc_debug_phase_cycles(out, 1)
mpc.reset(); mpc.solve_QP(None, traj)
lib.cmpc_debug_phase_cycles(out, 1)
cyc = np.array(out[:16]); cnt = np.array(out[16:])
print(f"B={B} stress={stress} mode={mode} kernel_ms={mpc.kernel_ms:.3f}  total cycles/QP = {cyc.sum() / B:.0f}")
for i, nm in enumerate(NAMES):
    if cnt[i]:
        print(f"  {nm:16s} visits {int(cnt[i]):7d}  cycles/visit {cyc[i] / cnt[i]:9.0f}  cycles/QP {cyc[i] / B:9.0f}  share {100 * cyc[i] / cyc.sum():5.1f}%")
