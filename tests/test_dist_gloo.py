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


REGION_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r); sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    import crumble_b200 as cb
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    data, nr, nb = cb.simulate("C1", 0.1, 11, threads=1)                # the SAME contig on every rank (strong scaling)
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); batch = bb.finish()
    shards, end = cb.plan_region_shards(batch, world)
    sh = shards[rank]
    sub, keep = cb.sub_batch(batch, sh["h0"], sh["r1"])
    assert sub.n_reads == sh["r1"] - sh["h0"] and cb.aligned_bases(sub) > 0
    n = int(batch.n_reads)
    done = np.zeros(n, dtype=bool)
    for k in range(rank):
        f = cb.shard_final_mask(batch, shards[k], end, done); done[np.arange(shards[k]["h0"], shards[k]["r1"])[f]] = True
    fin = cb.shard_final_mask(batch, sh, end, done)
    mine = np.zeros(n, dtype=np.int32); mine[np.arange(sh["h0"], sh["r1"])[fin]] = 1
    t = torch.from_numpy(mine); dist.all_reduce(t, op=dist.ReduceOp.SUM)
    # the 128-byte state of rank r travels to rank r+1 in position order (bench.py --region-shards does it with NCCL)
    blob = torch.full((cb.api.CARRY_BYTES,), rank + 1, dtype=torch.uint8)
    got = []
    for k in range(1, world):
        b = blob.clone(); dist.broadcast(b, src=k - 1)
        if rank == k: got.append(int(b[0]))
    if rank > 0: assert got == [rank]
    if rank == 0:
        assert int(t.min()) == 1 and int(t.max()) == 1, "every record must turn final in exactly one shard"
        print("REGION_OK", [s["r0"] - s["h0"] for s in shards])
    dist.barrier(); dist.destroy_process_group()
""") % (str(ROOT), str(ROOT / "tests"))


def test_two_rank_region_shards_gloo(tmp_path):
    """host side of the strong-scaling path: both ranks plan the same region shards of one contig, each owns one, every record
    turns final in exactly one of them, and the carry blob travels in position order"""
    script = tmp_path / "r.py"
    script.write_text(REGION_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0 and "REGION_OK" in r.stdout, r.stdout[-3000:]
