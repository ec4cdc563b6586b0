#!/usr/bin/env python
"""tools/sass_hot.py REPORT.ncu-rep [N] — SASS instructions of the captured kernel ranked by executed count, with region totals."""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ie = h.index("Instructions Executed"); si = h.index("Source"); sa = h.index("# Samples"); at = h.index("Avg. Threads Executed")
data = []
for r in rows[hi + 1:]:
    if len(r) <= ie: continue
    try: data.append((int(r[ie]), int(r[sa] or 0), float(r[at] or 0), r[si], r[0]))
    except ValueError: pass
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("total warp-instructions", tot, "samples", ts, "sass lines", len(data))
# contiguous regions of similar execution count
print("--- regions (runs of instructions with count within 2x of the run's first) ---")
i = 0
while i < len(data):
    j = i; s = 0; sm = 0
    while j < len(data) and data[i][0] > 0 and data[j][0] * 2 >= data[i][0] and data[j][0] <= data[i][0] * 2:
        s += data[j][0]; sm += data[j][1]; j += 1
    if j == i: j = i + 1
    elif s > tot * 0.01:
        print(f"lines {i:5d}-{j:5d} n={j-i:4d} instr={s:12d} ({s/tot:.3f}) samples={sm/ts:.3f} per-line={data[i][0]} first: {data[i][3][:70]}")
    i = j
print("--- top lines ---")
for d in sorted(data, reverse=True)[:N]:
    print(f"{d[0]:12d} {d[1]:6d} {d[2]:5.1f} {d[3][:100]}")
