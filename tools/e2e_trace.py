#!/usr/bin/env python
"""tools/e2e_trace.py [scale] : host-side timeline of cg_process (CG_TRACE) on C2, for several upload-chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import crumble_b200 as cb
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
data, nr, nb = cb.simulate("C2", scale, seed=100)
bb = cb.BatchBuilder(pinned=True); bb.add_bam_stream(data); del data
batch = bb.finish()
g = cb.Crumble(cb.default_params(9), device=0)
qout = torch.empty(max(int(batch.qual_bytes), 1), dtype=torch.uint8, pin_memory=True).numpy()
g.process(batch, pinned_out=qout)
for mb in (96, 64, 48, 32):
    g.set_chunk_bytes(mb << 20)
    g.process(batch, pinned_out=qout)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); g.process(batch, pinned_out=qout); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    t = g.timers()
    print(f"chunk {mb} MB: e2e {min(ts):.2f} / {np.mean(ts):.2f} ms  h2d span {t['h2d']:.2f}  d2h span {t['d2h']:.2f}", flush=True)
g.set_chunk_bytes(96 << 20)
os.environ["CG_TRACE"] = "1"
g.process(batch, pinned_out=qout)
