"""Per-SASS-instruction stall reasons of a source-line range: python tools/ncu_region.py rep lib kernel file lo hi"""
import csv, io, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_lines
rep, lib, kernel, fname, lo, hi = sys.argv[1:7]
lo, hi = int(lo), int(hi)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi_ = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi_]
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = rows[hi_ + 1:]
base = min(int(r[ia], 16) for r in data if len(r) > ia)
lines = sass_lines(lib, kernel)
tot = {}
n = 0
for r in data:
    if len(r) <= max(stall_cols):
        continue
    a = int(r[ia], 16) - base
    (f, l), _ = lines.get(a, (("?", 0), ""))
    if f != fname or not (lo <= l <= hi):
        continue
    samp = int(r[isamp] or 0)
    st = {hdr[i][6:]: int(r[i] or 0) for i in stall_cols if int(r[i] or 0)}
    for k, v in st.items():
        tot[k] = tot.get(k, 0) + v
    n += samp
    if samp >= int(sys.argv[7]) if len(sys.argv) > 7 else samp >= 20:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{a:6x} L{l:4d} samp {samp:5d} ex {int(r[iex] or 0):8d}  {r[isrc][:46]:46s} {top}")
print("region samples", n, sorted(tot.items(), key=lambda kv: -kv[1]))
