"""Multi-GPU plumbing: the robot batch is sharded across ranks, one process per GPU, with NO per-step
communication (every robot's QP is independent -- SURVEY.md section 8e).  ``torch.distributed`` is used
once, at the end of a run, to gather a fixed-size statistics record per rank (NCCL on GPUs, gloo in the
CPU tests).  Nothing here touches the solver."""
import os

import numpy as np

STAT_FIELDS = ("qps_count", "elapsed_ms", "solved", "max_iter", "inaccurate", "failed",
               "path_unconstrained", "path_active_set", "path_admm", "path_admm_polish",
               "as_iters_sum", "admm_iters_sum", "n_free_sum", "r_prim_max", "r_dual_max", "flops", "path_riccati", "flops_route", "path_wrench_as")


def env_rank_world():
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def shard_range(B, rank, world):
    """Contiguous slice [lo, hi) of ``ceil(B/world)`` robots owned by ``rank``."""
    per = -(-B // world)
    return min(rank * per, B), min((rank + 1) * per, B)


def local_stats(status, iters, stats, elapsed_ms, flops=0.0, flops_route=0.0):
    """Fixed-size float64 record (len(STAT_FIELDS)) from one rank's per-QP outputs (NumPy arrays)."""
    status = np.asarray(status)
    stats = np.asarray(stats)
    path = stats[:, 7].astype(np.int64) if len(stats) else np.zeros(0, np.int64)
    rec = np.zeros(len(STAT_FIELDS))
    rec[0] = len(status)
    rec[1] = elapsed_ms
    rec[2] = (status == 1).sum()
    rec[3] = (status == -2).sum()
    rec[4] = (status == 2).sum()
    rec[5] = ((status != 1) & (status != 2) & (status != -2)).sum()
    for p in range(4):
        rec[6 + p] = (path == p).sum()
    if len(stats):
        rec[10] = stats[:, 6].sum()
        rec[11] = np.asarray(iters).sum()
        rec[12] = stats[:, 3].sum()
        rec[13] = stats[:, 0].max()
        rec[14] = stats[:, 1].max()
    rec[15] = flops
    rec[16] = (path == 4).sum()          # finished by the Riccati pre-pass (no active constraint)
    rec[17] = flops_route
    rec[18] = (path == 5).sum()          # active set settled by Riccati sweeps (wrench-space PDAS kernel)
    return rec


def gather_stats(rec, device=None):
    """all_gather the per-rank records -> (world, len(STAT_FIELDS)) array on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(rec)[None, :]
    t = torch.tensor(np.asarray(rec), dtype=torch.float64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def reduce_stats(all_recs):
    """Whole-job summary from the gathered records: sums for counts, max for time and residuals."""
    a = np.asarray(all_recs)
    out = {}
    for i, name in enumerate(STAT_FIELDS):
        out[name] = float(a[:, i].max()) if name in ("elapsed_ms", "r_prim_max", "r_dual_max") else float(a[:, i].sum())
    return out
