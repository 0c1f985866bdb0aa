"""Per-kernel totals of the last complete gradient pass in an ncu launch list (gpu__time_duration.sum CSV)."""
import csv, re, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
H = rows[0]; ki = H.index("Kernel Name"); vi = H.index("Metric Value"); gi = H.index("Grid Size")
names = [(re.sub(r"\(.*", "", r[ki]).replace("void ", ""), r[gi], int(r[vi].replace(",", ""))) for r in rows[1:]]
idx = [i for i, n in enumerate(names) if n[0].startswith("k_wf_trace_primary")]
a, b = (idx[-2], idx[-1]) if len(idx) > 1 else (idx[-1], len(names))
tot = collections.OrderedDict()
for n, g, t in names[a:b]:
    if n.startswith("cub::") or n.startswith("at::"): n = "(library) " + n.split("<")[0]
    tot.setdefault(n, [0, 0]); tot[n][0] += 1; tot[n][1] += t
s = sum(v[1] for v in tot.values())
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]): print("%-60s %3d launches %8.3f ms  %.3f" % (n[:60], c, t / 1e6, t / s))
print("sum %.3f ms over %d launches" % (s / 1e6, b - a))
for n, g, t in names[a:b]: print("   ", n[:50], g, "%.3f" % (t / 1e6))
