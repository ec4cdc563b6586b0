#!/usr/bin/env python
"""tools/launch_summary.py LAUNCHES.csv OUT.tsv "command" — per-kernel totals and shares of an ncu launch list
(gpu__time_duration.sum, --clock-control none)."""
import csv, sys, re, collections
src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*$", "", r[ki]).strip()
    t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += float(r[vi].replace(",", "")) / 1e6
allms = sum(v[1] for v in tot.values())
with open(dst, "w") as f:
    f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none) of `{cmd}`, {len(rows)-1} launches\n")
    f.write("# cold-cache, serialised: compare SHARES not absolutes\nkernel\tlaunches\ttotal_ms\tshare\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k}\t{v[0]}\t{v[1]:.3f}\t{v[1]/allms:.3f}\n")
print(open(dst).read()[:1500])
