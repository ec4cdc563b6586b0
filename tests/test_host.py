"""CPU tests of the host side: option surface, hts_lite I/O, the record batcher, and that the
C-ABI library loads, exports every symbol include/crumble_gpu.h declares, and fails loudly
(no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import crumble_b200 as cb
from util import EMU_BIN, PORT_BIN, REF_BIN, ROOT

GDIR = ROOT / "tests" / "golden"


def test_library_exports_every_declared_symbol():
    hdr = open(ROOT / "include" / "crumble_gpu.h").read()
    declared = set(re.findall(r"\b(cg[bm]?_[a-z_0-9]+)\s*\(", hdr)) - {"cg_ctx"}
    lib = cb.load_lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(cb.EXPORTS) <= declared
    assert lib.cg_abi_version() == 3


def test_no_device_fails_loudly():
    lib = cb.load_lib()
    if lib.cg_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cb.CrumbleError, match="no usable CUDA device"):
        cb.Crumble()
    err = C.c_int(0)
    p = cb.default_params()
    assert not lib.cg_create(C.byref(p), 0, C.byref(err)) and err.value == -1
    # the command line must fail too, not silently compute on the CPU
    r = subprocess.run([str(ROOT / "crumble_b200" / "lib" / "crumble_gpu"), "-z", str(GDIR / "edge_cases.sam"), "/dev/null"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Error while reducing file" in r.stderr


def test_product_does_not_reference_the_oracle():
    """the shipped library must not link, call or embed anything under oracle/ or tests/emu/"""
    out = subprocess.run(["nm", "-D", "--undefined-only", str(cb.lib_path())], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in out and "emu" not in out
    import re as _re
    for f in (ROOT / "crumble_b200" / "csrc").rglob("*"):
        if f.is_file() and f.suffix in (".c", ".cpp", ".cu", ".h", "") and "build" not in f.parts:
            txt = f.read_text(errors="ignore")
            incs = _re.findall(r'#include\s+[<"]([^>"]+)[>"]', txt)
            assert not [i for i in incs if "oracle" in i or "emu" in i], f
            if f.name == "Makefile":
                assert "oracle" not in txt and "tests/" not in txt


LEVEL_ARGS = [["-9"], ["-8"], ["-7"], ["-5"], ["-3"], ["-1"], ["-1", "-B", "-u45", "-l3", "-c20"], ["-9", "-Q60", "-D100", "-X1.2", "-m3"],
              ["-i1.5,4", "-s0.5,2", "-q50", "-d60", "-x2.5"], ["-3", "-9"], ["-P5", "-C0.3", "-M0.4", "-Z0.2", "-V0.1", "-p4", "-L0"]]


@pytest.mark.skipif(not REF_BIN.exists(), reason="oracle/_ref not built")
@pytest.mark.parametrize("args", LEVEL_ARGS, ids=lambda a: "".join(a))
def test_option_parser_matches_reference_report(args):
    """our getopt loop + presets print the same -v parameter block as the reference main() (option order matters)."""
    sam = str(GDIR / "edge_cases.sam")
    a = subprocess.run([str(REF_BIN), "-z", "-v"] + args + [sam, "mem:x"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    b = subprocess.run([str(PORT_BIN), "-z", "-v"] + args + [sam, "mem:x"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert a.returncode == 0 and b.returncode == 0
    assert a.stdout == b.stdout
    pa = [l for l in a.stderr.splitlines() if " = " in l and not l.startswith(("A", "B", "Col", "Low_", "Clip_", "Ins_", "indel_ov_perc", "count_"))]
    pb = [l for l in b.stderr.splitlines() if " = " in l and not l.startswith(("A", "B", "Col", "Low_", "Clip_", "Ins_", "indel_ov_perc", "count_"))]
    assert pa == pb and len(pa) == 5


def test_python_level_presets_match_c():
    p9, p1 = cb.default_params(9), cb.default_params(1)
    assert (p9.pblock, p9.min_qual_B, p9.min_indel_B, p9.iSTR_add, p9.over_depth) == (8, 70, 125, 2, 999.0)
    assert (p1.pblock, p1.min_qual_B, p1.min_mqual, p1.sSTR_add, p1.iSTR_mul, p1.over_depth) == (0, 75, 5, 5, 2.0, 3.0)
    d = cb.default_params()
    assert (d.qlow, d.qcutoff, d.qhigh, d.qcap, d.clip_perc, d.region_tid) == (5, 25, 40, 60, 0.2, -1)


def test_hts_lite_sam_bam_roundtrip():
    """SAM -> BGZF BAM -> SAM and SAM -> raw BAM -> SAM through the identity option set (-p0 -Q0 -L0)."""
    sam = str(GDIR / "edge_cases.sam")
    with tempfile.TemporaryDirectory() as td:
        s1, bam, ubam, s2, s3 = (os.path.join(td, x) for x in ("a.sam", "a.bam", "a.ubam", "b.sam", "c.sam"))
        ident = ["-z", "-p0", "-Q0", "-L0"]
        for cmd in ([sam, s1], [sam, bam], [bam, s2], [sam, ubam], [ubam, s3]):
            subprocess.run([str(PORT_BIN)] + ident + cmd, check=True)
        body = lambda f: [l for l in open(f) if not l.startswith("@")]
        orig = body(sam)
        assert body(s1) == orig and body(s2) == orig and body(s3) == orig
        import gzip
        assert gzip.open(bam).read()[:4] == b"BAM\1" and open(ubam, "rb").read(4) == b"BAM\1"
        # @PG line is added without -z
        subprocess.run([str(PORT_BIN), "-p0", "-Q0", "-L0", sam, s1], check=True)
        # (the oracle CLI does not add @PG; the product CLI does — checked on the GPU box)


def test_batcher_layout_and_accounting():
    data, nr, nb = cb.simulate("tiny", 0.2, 9, threads=1)
    bb = cb.BatchBuilder(pinned=False)
    bb.add_bam_stream(data)
    b = bb.finish()
    assert b.n_reads == nr
    off, ln = bb.offsets(), bb.lengths()
    assert (off % 8 == 0).all() and (np.diff(off) >= ln[:-1]).all()
    assert cb.aligned_bases(b) == nb
    ncig = np.ctypeslib.as_array(b.n_cigar, shape=(nr,)).astype(np.int64)
    flag = np.ctypeslib.as_array(b.flag, shape=(nr,)); tid = np.ctypeslib.as_array(b.tid, shape=(nr,))
    inp = (tid >= 0) & ((flag & 4) == 0) & (ncig > 0)
    exp = int((((ln[inp] + 1) // 2) + 2 * ln[inp].astype(np.int64) + 4 * ncig[inp] + 16).sum())
    assert cb.algorithmic_bytes(b) == exp          # 395 B per 150-bp single-op read
    one = cb.BatchBuilder(pinned=False)
    one.add(0, 5, 0, 60, np.array([150 << 4], np.uint32), np.zeros(75, np.uint8), np.full(150, 30, np.uint8))
    assert cb.algorithmic_bytes(one.finish()) == 395


def test_batcher_rejects_unsorted():
    bb = cb.BatchBuilder(pinned=False)
    c = np.array([10 << 4], np.uint32); s = np.zeros(5, np.uint8); q = np.full(10, 30, np.uint8)
    bb.add(0, 100, 0, 60, c, s, q)
    bb.add(0, 50, 0, 60, c, s, q)
    with pytest.raises(cb.CrumbleError, match="sorted"):
        bb.finish()


def test_hts_lite_streaming_bgzf_multithreaded():
    """A BAM larger than the reader's chunk and the writer's parallel round: raw BAM -> BGZF BAM (4 deflate threads) must
    gunzip to the single-threaded raw stream, and reading it back in chunks (inflate threads) must give the same records."""
    import gzip
    import numpy as np
    import crumble_b200 as cb
    data, nr, nb = cb.simulate("C1", 1.0, seed=21, threads=2)
    assert data.nbytes > 40 << 20
    ident = ["-z", "-p0", "-Q0", "-L0"]
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=tmpdir) as td:
        raw, bam_mt, bam_st, back = (os.path.join(td, x) for x in ("in.ubam", "mt.bam", "st.bam", "back.ubam"))
        data.tofile(raw)
        subprocess.run([str(PORT_BIN)] + ident + ["-O", "bam,nthreads=4", raw, bam_mt], check=True)
        subprocess.run([str(PORT_BIN)] + ident + ["-O", "bam,nthreads=1", raw, bam_st], check=True, env=dict(os.environ, HTS_LITE_THREADS="1"))
        assert os.path.getsize(bam_mt) > 9 << 20                      # more than one 8 MiB read chunk
        a, b = gzip.open(bam_mt).read(), gzip.open(bam_st).read()
        assert a == b and a == data.tobytes()
        assert open(bam_mt, "rb").read() == open(bam_st, "rb").read()  # block boundaries do not depend on the thread count
        subprocess.run([str(PORT_BIN)] + ident + ["-I", "bam,nthreads=3", "-O", "bam,raw", bam_mt, back], check=True)
        assert np.array_equal(np.fromfile(back, dtype=np.uint8), data)


def test_region_shard_plan_invariants():
    """the host-side shard planner: windows are monotone, halos hold every earlier record that reaches beyond the shard's
    first column, every record turns final in exactly one shard"""
    import numpy as np
    import crumble_b200 as cb
    for preset, scale, seed, k in (("tiny", 1.0, 3, 7), ("C1", 0.1, 11, 4), ("C4", 0.02, 4, 3)):
        data, nr, nb = cb.simulate(preset, scale, seed=seed)
        bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); batch = bb.finish()
        shards, end = cb.plan_region_shards(batch, k)
        n = int(batch.n_reads)
        pos, tid = bb.positions(), np.ctypeslib.as_array(batch.tid, shape=(n,))
        done = np.zeros(n, dtype=bool)
        prev = None
        for sh in shards:
            assert sh["h0"] <= sh["r0"] < sh["r1"]
            if sh["first"] == 2:
                assert prev is not None and (sh["lo_tid"], sh["lo_pos"], sh["cnt_pos"]) == (prev["hi_tid"], prev["next_lo_pos"], prev["hi_pos"])
                assert sh["lo_pos"] <= sh["cnt_pos"]
                reach = np.nonzero((tid[:sh["r0"]] == sh["lo_tid"]) & (end[:sh["r0"]] > sh["lo_pos"]))[0]
                assert reach.size == 0 or reach[0] >= sh["h0"]
                assert not np.any(~done[: sh["h0"]])                       # nothing before the halo is still open
            fin = cb.shard_final_mask(batch, sh, end, done)
            idx = np.arange(sh["h0"], sh["r1"])[fin]
            assert not done[idx].any()
            done[idx] = True
            prev = sh
        assert done.all()
        bb.close()


def test_streaming_driver_error_paths():
    """transcode_gpu's reader / writer threads must surface failures as the reference does (exit 1, "Error while reducing file"):
    a truncated BAM stream, input that stops being coordinate sorted across a cut, an output that cannot be written."""
    import numpy as np
    import crumble_b200 as cb
    from util import EMU_BIN
    data, nr, nb = cb.simulate("tiny", 0.3, seed=5)
    with tempfile.TemporaryDirectory() as td:
        good, cut, out = (os.path.join(td, x) for x in ("in.ubam", "cut.ubam", "out.ubam"))
        data.tofile(good); data[: data.size // 2 + 7].tofile(cut)
        env = dict(os.environ, CRUMBLE_BATCH_READS="500")
        r = subprocess.run([str(EMU_BIN), "-z", "-9", "-O", "bam,raw", good, out], env=env, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([str(EMU_BIN), "-z", "-9", "-O", "bam,raw", cut, out], env=env, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "Error while reducing file" in r.stderr
        # unsorted across a cut: the second half of the records first
        sam = [l for l in open(GDIR / "edge_cases.sam")]
        hdr = [l for l in sam if l.startswith("@")]; body = [l for l in sam if not l.startswith("@")]
        mapped = [l for l in body if l.split("\t")[2] == "chrA"]
        bad = os.path.join(td, "bad.sam")
        open(bad, "w").write("".join(hdr + mapped[len(mapped) // 2:] + mapped[: len(mapped) // 2]))
        r = subprocess.run([str(EMU_BIN), "-z", "-9", bad, out], env=dict(os.environ, CRUMBLE_BATCH_READS="3"), stderr=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "Error while reducing file" in r.stderr
        r = subprocess.run([str(EMU_BIN), "-z", "-9", good, "/nonexistent_dir/out.bam"], env=env, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 1


@pytest.mark.parametrize("preset,scale,bits", [("C2", 1 / 256, 2), ("C1", 0.05, 0), ("tiny", 1.0, 0)])
def test_batcher_compact_planes_decode_back(preset, scale, bits):
    """cgb_pack: 2-bit bases + exception list and dictionary-coded qualities decode back to exactly the 4-bit / 8-bit arrays the
    kernels work on (checked on the host with numpy; the device expansion is checked against the same arrays in the GPU suite)"""
    data, nr, nb = cb.simulate(preset, scale, seed=21, threads=2)
    bb = cb.BatchBuilder(pinned=False)
    bb.add_bam_stream(data)
    b = bb.finish(pack=True, threads=3)
    n = int(b.qual_bytes)
    assert b.seq2 and int(b.seq2_bytes) == n // 4 and b.qual_bits == bits
    off, ln = bb.offsets(), bb.lengths()
    valid = np.zeros(n + 1, np.int32); np.add.at(valid, off, 1); np.add.at(valid, off + ln, -1); valid = np.cumsum(valid[:-1]) > 0
    seq = np.ctypeslib.as_array(b.seq, shape=(int(b.seq_bytes),))
    nib = np.empty(n, np.uint8); nib[0::2] = seq[: n // 2] >> 4; nib[1::2] = seq[: n // 2] & 15
    s2 = np.ctypeslib.as_array(b.seq2, shape=(n // 4,))
    code = np.empty(n, np.uint8)
    for k in range(4):
        code[k::4] = (s2 >> (2 * k)) & 3
    dec = (1 << code).astype(np.uint8)
    ne = int(b.n_seq_exc)
    if ne:
        exc = np.ctypeslib.as_array(b.seq_exc, shape=(ne,))
        assert np.all(np.diff(exc >> np.uint64(4)).astype(np.int64) > 0)
        dec[(exc >> np.uint64(4)).astype(np.int64)] = (exc & np.uint64(15)).astype(np.uint8)
    assert np.array_equal(dec[valid], nib[valid])
    assert ne == int((~np.isin(nib[valid], [1, 2, 4, 8])).sum())
    if bits:
        qp = np.ctypeslib.as_array(b.qualp, shape=(int(b.qualp_bytes),))
        per = 8 // bits
        qc = np.empty(n, np.uint8)
        for k in range(per):
            qc[k::per] = (qp >> (bits * k)) & ((1 << bits) - 1)
        dq = np.array(list(b.qual_dict), np.uint8)[qc]
        assert np.array_equal(dq[valid], bb.qual()[valid])
    bb.close()


@pytest.mark.parametrize("preset,scale", [("tiny", 1.0), ("C1", 0.2), ("C4", 0.1)])
def test_batcher_record_planes_decode_back(preset, scale):
    """cgb_pack: the compact planes of the per-record arrays (tid runs, position deltas + listed positions, read-length codes, CIGAR counts
    with the implied single-M form, the remaining CIGAR operations) decode back to tid / pos / l_qseq / n_cigar / cigar."""
    data, nr, nb = cb.simulate(preset, scale, seed=33, threads=2)
    bb = cb.BatchBuilder(pinned=False)
    bb.add_bam_stream(data)
    b = bb.finish(pack=True, threads=2)
    n = int(b.n_reads)
    assert b.meta_planes == 1 and n > 0
    A = lambda p, k: np.ctypeslib.as_array(p, shape=(int(k),))
    tid, pos, lq, nc = A(b.tid, n), A(b.pos, n), A(b.l_qseq, n), A(b.n_cigar, n)
    runs = A(b.tid_runs, b.n_tid_runs); idx = (runs >> np.uint64(32)).astype(np.int64); val = (runs & np.uint64(0xffffffff)).astype(np.uint32).astype(np.int32)
    assert idx[0] == 0 and np.all(np.diff(idx) > 0)
    assert np.array_equal(val[np.searchsorted(idx, np.arange(n), side="right") - 1], tid)
    d8 = A(b.pos_d8, n).astype(np.int64); pa = A(b.pos_abs, b.n_pos_abs)
    ai = (pa >> np.uint64(32)).astype(np.int64); av = (pa & np.uint64(0xffffffff)).astype(np.uint32).astype(np.int32).astype(np.int64)
    assert ai[0] == 0 and np.all(np.diff(ai) > 0) and np.all(d8[ai] == 255) and int((d8 == 255).sum()) == ai.size
    S = np.cumsum(np.where(d8 == 255, 0, d8))
    k = np.searchsorted(ai, np.arange(n), side="right") - 1
    assert np.array_equal(av[k] + (S - S[ai][k]), pos.astype(np.int64))
    assert np.array_equal(np.array(list(b.lq_dict), np.int32)[A(b.lq8, n)], lq)
    nc8 = A(b.nc8, n).astype(np.int64)
    assert np.array_equal(np.where(nc8 == 255, 1, nc8), nc.astype(np.int64))
    cig = A(b.cigar, b.n_cigar_total); coff = A(b.cigar_off, n).astype(np.int64)
    cx = A(b.cigar_x, b.n_cigar_x) if b.n_cigar_x else np.zeros(0, np.uint32)
    xoff = np.cumsum(np.where(nc8 == 255, 0, nc8)) - np.where(nc8 == 255, 0, nc8)
    imp = nc8 == 255
    assert np.array_equal(cig[coff[imp]], (lq[imp].astype(np.uint32) << np.uint32(4)))
    for i in np.nonzero(~imp)[0][:2000]:
        assert np.array_equal(cig[coff[i]: coff[i] + nc8[i]], cx[xoff[i]: xoff[i] + nc8[i]])
    assert int(b.n_cigar_x) == int(np.where(imp, 0, nc8).sum())
    bb.close()


TAGS = {"plain": ["-9"], "t": ["-9", "-t", "NM,BD"], "T": ["-9", "-T", "MD,XX,ZB"], "efg": ["-9", "-e", "5", "-f", "20", "-g", "40"],
        "EFGt": ["-9", "-E", "7", "-F", "25", "-G", "45", "-t", "BI,BD,RG"], "all": ["-1", "-e", "1", "-f", "30", "-g", "2", "-E", "3", "-F", "10", "-G", "50", "-T", "BI"]}


@pytest.mark.parametrize("tag", sorted(TAGS))
def test_aux_tag_options_match_reference(tag):
    """-t / -T tag lists and -e -f -g / -E -F -G BD / BI binarisation (purge_tags, snp_score.c:989-1054, 2031-2054) through the
    product's host driver (here on the CPU emulation of the device): the WHOLE SAM output equals the reference's, record for record,
    for reads carrying i, Z, A, f, H and B-array tags in shuffled order"""
    from util import EMU_BIN
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "o.sam")
        r = subprocess.run([str(EMU_BIN), "-z"] + TAGS[tag] + [str(GDIR / "tags.sam"), out], stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        assert open(out).read() == open(GDIR / f"tags.{tag}.out.sam").read()


def _mailbox_rank(world, rank, key, steps, out):
    import random
    import sys
    import time
    sys.path.insert(0, str(ROOT))
    import bench
    mb = bench.Mailbox(world, rank, key)
    time.sleep(0.3 if rank else 0.0)                         # rank 0 creates the file
    mb.open()
    rnd = random.Random(rank)
    acc = 0
    for s in range(1, steps + 1):
        if rank and rnd.random() < 0.05:
            time.sleep(0.003 * rnd.random() * world)         # a slow step now and then: rank 0 (which waits for nobody) runs ahead
        got = int.from_bytes(mb.recv(s, rank - 1)[:8], "little") if rank else 0
        val = got + s * (rank + 1)
        acc += val
        if rank < world - 1:
            mb.send(s, val.to_bytes(8, "little") + bytes(120))
    out.put((rank, acc))


def test_bench_mailbox_survives_drifting_ranks():
    """bench.py's shared-memory mailbox (the carry of a region shard, rank to rank): ranks that drift several steps apart neither lose
    nor overwrite a message (4 processes, 400 steps, random stalls)."""
    import multiprocessing as mp
    world, steps, key = 4, 400, f"test{os.getpid()}"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_mailbox_rank, args=(world, r, key, steps, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=30)
    try:
        os.unlink(f"/dev/shm/crumble_mbox_{key}")
    except OSError:
        pass
    exp = {r: sum(s * sum(k + 1 for k in range(r + 1)) for s in range(1, steps + 1)) for r in range(world)}
    assert res == exp
