"""Compile the CUDA library in-tree for sm_100a (B200).  No JIT cache, no torch extension machinery:
one nvcc command, the resulting ``libcmpc.so`` sits next to this file and travels with the repo."""
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "cmpc.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "cmpc_core.cuh"), os.path.join(HERE, "csrc", "cmpc_fast.cuh"),
        os.path.join(HERE, "csrc", "cmpc_riccati.cuh"), os.path.join(HERE, "csrc", "cmpc_riccati2.cuh"), os.path.join(HERE, "csrc", "cmpc_traj.cuh"), os.path.join(HERE, "csrc", "cmpc_wrench.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "cmpc.h")]
LIB = os.path.join(HERE, "libcmpc.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def find_nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build libcmpc.so")


STAMP = os.path.join(HERE, "libcmpc.so.stamp")


def source_hash():
    """sha256 over the sources and the compile flags: what the stamp next to libcmpc.so records."""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for d in DEPS:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale():
    """The library is current iff its stamp holds the hash of today's sources (file times are not trusted: a
    prebuilt libcmpc.so travels with repository snapshots and may be newer than a fresh checkout's sources)."""
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build_timing():
    """Tools only: the same source with per-phase clock64 accounting -> libcmpc_timing.so."""
    out = os.path.join(HERE, "libcmpc_timing.so")
    cmd = [find_nvcc()] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DCMPC_PHASE_TIMING", "-o", out, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stderr)
        raise RuntimeError("nvcc failed building libcmpc_timing.so")
    return out


def build(force=False, verbose=False):
    """Build ``libcmpc.so`` if missing or older than its sources.  Returns the path."""
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(" ".join(cmd))
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libcmpc.so")
    with open(os.path.join(HERE, "csrc", "ptxas_info.txt"), "w") as f:
        f.write(res.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    import sys
    if "--timing" in sys.argv:
        print(build_timing())
    else:
        print(build(force=True, verbose=True))
