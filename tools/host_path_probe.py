"""Where the time of the host-buffer entry (cmpc_solve_host) goes: pure H2D / D2H copy times of one batch against the
chunked call at several chunk sizes (development aid; python tools/host_path_probe.py [B])."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convex_mpc_b200 import records  # noqa: E402
from convex_mpc_b200.centroidal_mpc import BatchedComTraj, CentroidalMPC  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = 16
rec = records.random_records(B, seed=65536)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hb = [pin(rec.x0), pin(rec.x_ref), pin(rec.r_foot), pin(rec.I_world), pin(rec.mass), pin(rec.t0)]
out = (torch.empty(B, 12 * N, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory(),
       torch.empty(B, dtype=torch.int32).pin_memory())
dev = [t.cuda() for t in hb]
du = torch.empty(B, 12 * N, dtype=torch.float64, device="cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3


def h2d():
    for d, h in zip(dev, hb):
        d.copy_(h, non_blocking=True)


def d2h():
    out[0].copy_(du, non_blocking=True)


s2 = torch.cuda.Stream()


def both():
    h2d()
    with torch.cuda.stream(s2):
        d2h()


res = {"B": B, "h2d_ms": timeit(h2d), "d2h_ms": timeit(d2h), "h2d+d2h_concurrent_ms": timeit(both)}
res["h2d_GBs"] = sum(t.numel() * t.element_size() for t in hb) / res["h2d_ms"] / 1e6
res["d2h_GBs"] = out[0].numel() * 8 / res["d2h_ms"] / 1e6
traj = BatchedComTraj.from_records(rec, device="cuda:0")
ms_ = 4 * (int(np.floor(rec.duty * N)) + 1)
for chunk in (0, 4096, 8192, 16384, 32768, 65536):
    if chunk:
        os.environ["CMPC_HOST_CHUNK"] = str(chunk)
    m = CentroidalMPC(None, traj, verbose=False, max_stance=ms_, max_batch=B)

    def step():
        m._warm_host = 0
        m.solve_host(*hb, rec.dt, rec.gait_hz, rec.duty, out=out)
    res[f"solve_host_chunk_{chunk or 'default'}_ms"] = timeit(step)
    del m
os.environ.pop("CMPC_HOST_CHUNK", None)
for sch in ("8192,24576,24576,8192", "8192,16384,16384,12288,8192,4096", "8192,20480,20480,8192,4096,4096", "12288,20480,16384,8192,4096,4096",
            "16384,16384,16384,8192,4096,4096", "8192,24576,16384,8192,4096,4096", "4096,12288,20480,16384,8192,4096"):
    os.environ["CMPC_HOST_CHUNKS"] = sch
    m = CentroidalMPC(None, traj, verbose=False, max_stance=ms_, max_batch=B)

    def step():
        m._warm_host = 0
        m.solve_host(*hb, rec.dt, rec.gait_hz, rec.duty, out=out)
    res[f"solve_host_sched_{sch}_ms"] = timeit(step)
    del m
print(json.dumps(res, indent=1))
