"""Oracle: trot contact schedule.  Restates reference ``convex_mpc/gait.py:8,13-19,26-37``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np

# gait.py:8 -- leg order FL, FR, RL, RR (diagonal pairs in phase: trot)
PHASE_OFFSET = (0.5, 0.0, 0.0, 0.5)


def gait_period(frequency_hz):
    """gait.py:17 -- the period is the floating-point reciprocal of the frequency."""
    return 1 / frequency_hz


def contact_table(t0, dt, N, frequency_hz=3, duty=0.6, phase_offset=PHASE_OFFSET):
    """(4, N) int32 stance table sampled at step mid-points (gait.py:26-37).

    The arithmetic order matters for bit-exactness and is kept literally:
    ``t_k = (t0 + k*dt) + dt/2``; ``phase = fmod(offset + t_k / T, 1)``; ``stance = phase < duty``.
    """
    T = gait_period(frequency_hz)
    k = np.arange(N)
    t = t0 + k * dt            # gait.py:29
    t = t + dt / 2             # gait.py:30
    off = np.asarray(phase_offset, dtype=np.float64).reshape(4)
    ph = np.mod(off[:, None] + t[None, :] / T, 1.0)   # gait.py:33
    return (ph < duty).astype(np.int32)                # gait.py:36


def pack_mask(table):
    """Pack a (..., 4, N) 0/1 table into 64-bit words, bit index ``leg*N + k`` (N <= 16 -> 1 word).

    This is the wire format of the C-ABI (``include/cmpc.h``); N in {16, 32, 48} -> 1/2/3 words.
    """
    table = np.asarray(table)
    N = table.shape[-1]
    flat = table.reshape(table.shape[:-2] + (4 * N,)).astype(np.uint64)
    W = (4 * N + 63) // 64
    out = np.zeros(table.shape[:-2] + (W,), dtype=np.uint64)
    for b in range(4 * N):
        out[..., b // 64] |= flat[..., b] << np.uint64(b % 64)
    return out


def unpack_mask(words, N):
    words = np.asarray(words, dtype=np.uint64)
    lead = words.shape[:-1]
    out = np.zeros(lead + (4, N), dtype=np.int32)
    for leg in range(4):
        for k in range(N):
            b = leg * N + k
            out[..., leg, k] = ((words[..., b // 64] >> np.uint64(b % 64)) & np.uint64(1)).astype(np.int32)
    return out
