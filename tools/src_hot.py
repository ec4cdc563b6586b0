#!/usr/bin/env python
"""tools/src_hot.py REPORT.ncu-rep KERNEL_REGEX [N] — CUDA source lines of the captured kernel ranked by executed warp instructions
(ncu --page source --print-source cuda,sass: SASS rows carry the metrics and follow the source line they belong to)."""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]; N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
acc = collections.OrderedDict(); fname = None; h = None; line = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": h = r; ie = h.index("Instructions Executed"); sa = h.index("# Samples"); continue
    if h is None: continue
    if r[0] != "": line = (fname, int(r[0]), r[1]); acc.setdefault(line, [0, 0])
    if len(r) > ie and r[2] != "":
        try: acc[line][0] += int(r[ie] or 0); acc[line][1] += int(r[sa] or 0)
        except ValueError: pass
tot = sum(v[0] for v in acc.values()) or 1; ts = sum(v[1] for v in acc.values()) or 1
print("total warp-instructions", tot, "samples", ts)
byfile = collections.Counter()
for (f, l, s), v in acc.items(): byfile[f] += v[0]
print({f: round(v / tot, 3) for f, v in byfile.items()})
if "--order" in sys.argv:
    for (f, l, s), v in acc.items():
        if v[0] > tot * 0.002: print(f"{v[0]:12d} {v[0]/tot:6.3f} {v[1]/ts:6.3f} {f}:{l:<5d} {s.strip()[:120]}")
else:
    for (f, l, s), v in sorted(acc.items(), key=lambda kv: -kv[1][0])[:N]:
        print(f"{v[0]:12d} {v[0]/tot:6.3f} {v[1]/ts:6.3f} {f}:{l:<5d} {s.strip()[:120]}")
