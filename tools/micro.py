"""Print the library's micro-benchmarks (roofline denominators): python tools/micro.py"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import _lib
lib = _lib.load()
a, b, c, d = (ctypes.c_double() for _ in range(4))
_lib.check(lib.cmpc_microbench(0, ctypes.byref(a), ctypes.byref(b)))
_lib.check(lib.cmpc_microbench_dmma(0, ctypes.byref(c), ctypes.byref(d)))
lat = (ctypes.c_double * 4)()
_lib.check(lib.cmpc_microbench_latency(0, lat))
print(json.dumps({"latency_cycles": {"dfma": lat[0], "rsqrt_plus_dadd": lat[1], "shfl64": lat[2], "lds64_f2i": lat[3]}}))
print(json.dumps({"fp64_fma_tflops": a.value, "smem_gbs": b.value, "dmma_tflops": c.value, "dmma_latency_cycles": d.value}))
