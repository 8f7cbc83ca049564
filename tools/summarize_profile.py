"""Turn gpurun_out artefacts into the committed summaries under profiles/:
    python tools/summarize_profile.py <tag> <launches.csv> <prof.ncu-rep> <lib.so> <kernel> [bench.json]"""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep, lib, kernel = sys.argv[1:6]
bench = sys.argv[6] if len(sys.argv) > 6 else None
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr, agg = None, collections.defaultdict(list)
for r in rows:
    if r[0] == "ID": hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") == "gpu__time_duration.sum":
        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
        v = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
        agg[d["Kernel Name"].split("(")[0]].append(v)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, f"{tag}_launches_summary.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 200 over `python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline`\n")
    f.write("# per-launch times are cold-cache and serialised: compare shares, not absolutes\nkernel,launches,total_ms,mean_ms,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k},{len(v)},{sum(v):.3f},{sum(v)/len(v):.3f},{sum(v)/tot:.4f}\n")

# ---- raw metrics of the profiled kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
names, units = rr[0], rr[1]
ki = names.index("Kernel Name")
vals = next((r for r in rr[2:] if kernel in r[ki]), rr[2])      # the report may hold several kernels
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max"]
m = {}
for k in want:
    if k in names:
        i = names.index(k); m[k] = f"{vals[i]} {units[i]}"
# ---- stall reasons (whole kernel) from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel}"], capture_output=True, text=True).stdout
sr = list(csv.reader(io.StringIO(src)))
# a report with several kernels prints one block per kernel ("Kernel Name" row, header row, instructions): keep the
# first block of the kernel asked for
starts = [i for i, r in enumerate(sr) if r and r[0] == "Kernel Name"]
if starts:
    b0 = next((i for i in starts if kernel in sr[i][1]), starts[0])
    b1 = next((i for i in starts if i > b0), len(sr))
    sr = sr[b0:b1]
h = next(i for i, r in enumerate(sr) if r and r[0] == "Address")
sh = sr[h]
cols = [i for i, n in enumerate(sh) if n.startswith("stall_") and "Not Issued" not in n]
st = collections.Counter()
for r in sr[h + 1:]:
    if len(r) > max(cols):
        for i in cols: st[sh[i][6:]] += int(r[i] or 0)
ssum = sum(st.values())
with open(os.path.join(out, f"{tag}_{kernel}_ncu_summary.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on, one launch of {kernel} (tools/prof_case_prepass.py 65536 0.0 40 4: the bench's own size)\n")
    for k, v in m.items(): f.write(f"{k:80s} {v}\n")
    f.write("\n# warp stall reasons, all samples\n")
    for k, v in st.most_common(): f.write(f"stall_{k:24s} {v:8d} {100*v/ssum:5.1f}%\n")
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, lib, kernel, "30"], capture_output=True, text=True).stdout
open(os.path.join(out, f"{tag}_{kernel}_stalls_by_function.txt"), "w").write(lines)
if bench:
    line = [l for l in open(bench) if l.startswith("{")][-1]
    json.dump(json.loads(line), open(os.path.join(out, f"{tag}_bench.json"), "w"), indent=1)
print(json.dumps(m, indent=1))
print(st.most_common(6))
