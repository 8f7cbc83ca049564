"""Stall-reason shares, headline metrics and per-source-line sample shares of one kernel in an .ncu-rep:
    python tools/ncu_quick.py <rep> [top_lines]"""
import csv, collections, subprocess, io, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); n, u, v = r[0], r[1], r[2]
for k in ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
          'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']:
    if k in n: i = n.index(k); print(k, v[i], u[i])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address'); hdr = rows[h]
stall = [i for i, nm in enumerate(hdr) if nm.startswith('stall_') and 'Not Issued' not in nm]
c = collections.Counter(); body = rows[h + 1:]
for r in body:
    if len(r) > max(stall):
        for j in stall: c[hdr[j][6:]] += int(r[j] or 0)
t = sum(c.values()); print([(k, round(100 * v / t, 1)) for k, v in c.most_common(9)])
# coarse histogram of samples over the SASS address range (20 buckets): where in the code the time goes
tot = [int(r[2] or 0) for r in body if len(r) > 3]
nb = 20; L = len(tot)
print('instructions', L, ' samples by code position (5% buckets):', [round(100 * sum(tot[i * L // nb:(i + 1) * L // nb]) / max(sum(tot), 1), 1) for i in range(nb)])
