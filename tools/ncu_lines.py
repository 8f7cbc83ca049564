"""Attribute ncu warp-stall samples of one kernel to source lines / functions.

    python tools/ncu_lines.py report.ncu-rep libcmpc.so solve_kernel [topN]

ncu's CSV source page is per SASS instruction; this joins it with ``nvdisasm --print-line-info`` of
the same cubin and sums samples per source line and per enclosing function of csrc/cmpc_core.cuh.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    m = {}
    cur = None
    inside = False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        f = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if f:
            cur = (os.path.basename(f.group(1)), int(f.group(2)))
            continue
        a = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if a and cur:
            m[int(a.group(1), 16)] = (cur, a.group(2).strip())
    return m


def function_ranges(path):
    """(start_line, name) for each function-like definition in a source file."""
    out = []
    for i, ln in enumerate(open(path), 1):
        f = re.match(r"^(?:CMPC_HDN?|CMPC_DEV|__global__|static|inline|__device__)[^;]*?\b([A-Za-z_0-9]+)\s*\(", ln)
        if f and not ln.strip().startswith("//"):
            out.append((i, f.group(1)))
    return out


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if starts:      # several kernels in the report: keep the first block of the one asked for
        b0 = next((i for i in starts if kernel in rows[i][1]), starts[0])
        b1 = next((i for i in starts if i > b0), len(rows))
        rows = rows[b0:b1]
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    ibar = hdr.index("stall_barrier")
    data = [(int(r[ia], 16), int(r[isamp] or 0), int(r[iex] or 0), int(r[ibar] or 0)) for r in rows[hdr_i + 1:] if len(r) > ibar]
    base = min(a for a, *_ in data)
    lines = sass_lines(lib, kernel)
    per_line = defaultdict(lambda: [0, 0, 0])
    for a, s, ex, bar in data:
        key = lines.get(a - base, (("?", 0), ""))[0]
        per_line[key][0] += s
        per_line[key][1] += ex
        per_line[key][2] += bar
    total = sum(v[0] for v in per_line.values())
    total_ex = sum(v[1] for v in per_line.values())
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(here, "convex-mpc-unitree-go2_b200", "csrc")
    src = {f: os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".cu"))}
    fr = {k: function_ranges(v) for k, v in src.items()}
    per_fn = defaultdict(lambda: [0, 0, 0])
    for (f, l), v in per_line.items():
        name = "?"
        for start, nm in fr.get(f, []):
            if start <= l:
                name = nm
        for i in range(3):
            per_fn[f"{f}:{name}"][i] += v[i]
    print(f"kernel {kernel}: {total} samples, {total_ex} warp-instructions")
    print("\n== by function (samples %, of which barrier-stall %, instructions %) ==")
    for k, v in sorted(per_fn.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / total:6.2f}%  bar {100 * v[2] / max(v[0], 1):5.1f}%  inst {100 * v[1] / total_ex:6.2f}%  {k}")
    print("\n== by line ==")
    text = {k: open(v).read().splitlines() for k, v in src.items()}
    for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        code = text[f][l - 1].strip() if f in text and 0 < l <= len(text[f]) else ""
        print(f"{100 * v[0] / total:6.2f}%  bar {100 * v[2] / max(v[0], 1):5.1f}%  inst {100 * v[1] / total_ex:6.2f}%  {f}:{l}  {code[:90]}")


if __name__ == "__main__":
    main()
