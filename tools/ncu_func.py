"""Per-source-line samples of chosen functions: python tools/ncu_func.py rep lib kernel func1,func2 [top]"""
import csv, io, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_lines, function_ranges
rep, lib, kernel, funcs = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isamp, iex, ibar = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_barrier")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
data = [r for r in rows[h + 1:] if len(r) > max(stall_cols)]
base = min(int(r[ia], 16) for r in data)
lines = sass_lines(lib, kernel)
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(here, "convex-mpc-unitree-go2_b200", "csrc", "cmpc_fast.cuh")
src = open(path).read().splitlines()
fr = function_ranges(path)
def fn_of(l):
    name = "?"
    for start, nm in fr:
        if start <= l: name = nm
    return name
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot)
for fn in funcs.split(","):
    agg, st = {}, {}
    for r in data:
        a = int(r[ia], 16) - base
        (f, l), _ = lines.get(a, (("?", 0), ""))
        if f != "cmpc_fast.cuh" or fn_of(l) != fn: continue
        e = agg.setdefault(l, [0, 0, 0, 0]); e[0] += int(r[isamp] or 0); e[1] += int(r[iex] or 0); e[2] += int(r[ibar] or 0); e[3] += 1
        for i in stall_cols:
            v = int(r[i] or 0)
            if v: st[hdr[i][6:]] = st.get(hdr[i][6:], 0) + v
    n = sum(v[0] for v in agg.values())
    print(f"== {fn}: samples {n} ({100*n/tot:.1f}%) barrier {sum(v[2] for v in agg.values())} instr {sum(v[1] for v in agg.values())}  stalls {sorted(st.items(), key=lambda kv: -kv[1])[:6]}")
    for l, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"   L{l:4d} samp {v[0]:6d} bar {v[2]:6d} ex {v[1]:9d} sass {v[3]:3d}  {src[l-1].strip()[:84]}")
