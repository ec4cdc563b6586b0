"""world_size-2 gloo test of the multi-process plumbing bench.py uses for N>1: every rank owns
one contig shard, nothing is exchanged on the data path; only the timing (MAX) and the unit
counts (SUM) are reduced.  The per-shard work is done by the host emulation of the device
pipeline here (no GPU in this container); on the GPU box bench.py does the same with NCCL."""
import os
import socket
import subprocess
import sys
import textwrap

from util import ROOT

WORKER = textwrap.dedent("""
    import os, sys, hashlib
    sys.path.insert(0, %r); sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    import crumble_b200 as cb
    from util import run_oracle, valid_mask, EMU_BIN, PORT_BIN
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    data, nr, nb = cb.simulate("tiny", 0.3, 100 + rank, threads=1)      # one contig shard per rank (weak scaling)
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); bb.finish(); m = valid_mask(bb)
    a = run_oracle(data, ["-9"], binary=EMU_BIN, kind="emu")
    b = run_oracle(data, ["-9"], binary=PORT_BIN, kind="port")
    ok = np.array_equal(a["qual"][m], b["qual"][m]) and a["bed"] == b["bed"]
    t = torch.tensor([float(rank + 1)], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u = torch.tensor([float(nb), float(ok), float(a["counters"]["columns"])], dtype=torch.float64); dist.all_reduce(u, op=dist.ReduceOp.SUM)
    h = [None] * world
    dist.all_gather_object(h, hashlib.sha256(data.tobytes()).hexdigest())
    if rank == 0:
        assert t.item() == world and u[1].item() == world, (t, u)
        assert len(set(h)) == world, "ranks must own different shards"
        print("GLOO_OK", int(u[0].item()), int(u[2].item()))
    dist.barrier(); dist.destroy_process_group()
""") % (str(ROOT), str(ROOT / "tests"))


def test_two_rank_shards_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-3000:]
