"""CPU tests of the *device decomposition*: tests/emu/ compiles the same work-item bodies
that the sm_100a kernels run (crumble_b200/csrc/cg_pipeline.h) into a host emulation and
checks them against the golden vectors.  This is how the per-column -> sparse chain ->
per-read replay split is debugged where no GPU exists; the GPU tests repeat it on the device."""
import hashlib
import json

import pytest

import crumble_b200 as cb
from test_oracle import EDGE, GDIR, GOLD, run_cli_sam, sim
from util import EMU_BIN, run_oracle, valid_mask


@pytest.mark.parametrize("name", sorted(GOLD))
def test_emulated_pipeline_matches_golden(name):
    data = sim(name)
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); bb.finish()
    m = valid_mask(bb)
    for args, exp in GOLD[name]["runs"].items():
        r = run_oracle(data, args.split(), binary=EMU_BIN, kind="emu")
        assert hashlib.sha256(r["qual"][m].tobytes()).hexdigest() == exp["qual_sha256"], args
        assert r["bed"] == exp["bed"], args
        assert r["counters"] == exp["counters"], args


@pytest.mark.parametrize("tag", sorted(EDGE))
def test_emulated_pipeline_edge_cases(tag):
    quals, bed = run_cli_sam(EMU_BIN, EDGE[tag], GDIR / "edge_cases.sam")
    exp = [tuple(l.rstrip("\n").split("\t")) for l in open(GDIR / f"edge_cases.{tag}.qual.txt")]
    assert quals == exp
    assert bed == open(GDIR / f"edge_cases.{tag}.bed").read()


OPT = [(n, a) for n in sorted(GOLD) for a in sorted(GOLD[n].get("opt_runs", {}))]


@pytest.mark.parametrize("name,args", OPT, ids=lambda v: v.replace(" ", "") if isinstance(v, str) else None)
def test_emulated_pipeline_options(name, args):
    """-S, -k/-K/-y pbccs, -N and -R keep.bed (the options the device serves through the plain per-item bodies)
    against golden vectors made with the verbatim reference."""
    data = sim(name)
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); bb.finish()
    m = valid_mask(bb)
    exp = GOLD[name]["opt_runs"][args]
    argv = [str(GDIR / f"keep.{name}.bed") if x == "BED" else x for x in args.split()]
    r = run_oracle(data, argv, binary=EMU_BIN, kind="emu")
    assert hashlib.sha256(r["qual"][m].tobytes()).hexdigest() == exp["qual_sha256"]
    assert r["bed"] == exp["bed"]
    assert r["counters"] == exp["counters"]


# ---- chained calls: the host driver cuts the stream into region shards with a read halo (transcode_gpu.c, cg_process_window) ----
CHAIN = [("tiny", 1), ("tiny", 97), ("c1s", 1500), ("c2s", 4000), ("c4s", 333)]
CHAIN_ARGS = ["-9", "-1", "-3", "-5 -q30", "-1 -m10 -C0.05 -Z0.01", "-3 -i0.5,3 -s2.0,1"]


@pytest.mark.parametrize("name,batch", CHAIN, ids=lambda v: str(v))
def test_chained_calls_match_golden(name, batch):
    """The same golden vectors with the input cut every `batch` records: every cut leaves reads open, so the halo, the
    replayed columns and both carried states (keep-window chain, depth average at -1/-3/-5) are all exercised."""
    data = sim(name)
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); bb.finish()
    m = valid_mask(bb)
    for args in CHAIN_ARGS if batch > 1 else CHAIN_ARGS[:2]:
        exp = GOLD[name]["runs"][args]
        r = run_oracle(data, args.split(), binary=EMU_BIN, kind="emu", env_extra={"CRUMBLE_BATCH_READS": str(batch)})
        assert hashlib.sha256(r["qual"][m].tobytes()).hexdigest() == exp["qual_sha256"], args
        assert r["bed"] == exp["bed"], args
        assert r["counters"] == exp["counters"], args


@pytest.mark.parametrize("batch", [1, 5, 40])
def test_chained_calls_edge_cases(batch):
    """odd CIGARs (N skips, long reads, clips), FUNMAP-placed and unplaced reads, two contigs, across cuts"""
    for tag in ("l9", "l1B", "l5q30", "l3U35"):
        quals, bed = run_cli_sam(EMU_BIN, EDGE[tag], GDIR / "edge_cases.sam", env_extra={"CRUMBLE_BATCH_READS": str(batch)})
        exp = [tuple(l.rstrip("\n").split("\t")) for l in open(GDIR / f"edge_cases.{tag}.qual.txt")]
        assert quals == exp, tag
        assert bed == open(GDIR / f"edge_cases.{tag}.bed").read(), tag


def test_chained_calls_event_buffer_regrow():
    """more BED events in one call than the caller's buffer holds: the driver fetches them again (cg_download) instead of
    redoing the call, because the chain's state has already moved on"""
    data = sim("c1s")
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); bb.finish()
    m = valid_mask(bb)
    exp = GOLD["c1s"]["runs"]["-1"]
    assert exp["bed"].count("\n") > 8
    r = run_oracle(data, ["-1"], binary=EMU_BIN, kind="emu", env_extra={"CRUMBLE_BATCH_READS": "6000", "CRUMBLE_EVENTS_CAP": "1"})
    assert hashlib.sha256(r["qual"][m].tobytes()).hexdigest() == exp["qual_sha256"]
    assert r["bed"] == exp["bed"] and r["counters"] == exp["counters"]


def test_chained_calls_options():
    """-S, -k/-K, -N, -R across cuts (the plain per-item bodies)"""
    data = sim("tiny")
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); bb.finish()
    m = valid_mask(bb)
    for args, exp in GOLD["tiny"]["opt_runs"].items():
        argv = [str(GDIR / "keep.tiny.bed") if x == "BED" else x for x in args.split()]
        r = run_oracle(data, argv, binary=EMU_BIN, kind="emu", env_extra={"CRUMBLE_BATCH_READS": "211"})
        assert hashlib.sha256(r["qual"][m].tobytes()).hexdigest() == exp["qual_sha256"], args
        assert r["bed"] == exp["bed"], args
        assert r["counters"] == exp["counters"], args


SPARSE_ARGS = [["-9"], ["-1"], ["-3", "-P1.5"]]


@pytest.mark.parametrize("depth", [1.2, 3.0])
def test_chained_calls_sparse_coverage(depth):
    """1-3x coverage with placed-unmapped reads, cut after every record: coverage gaps, calls whose only new record never enters
    the pileup (no column at all), and at -1 / -P1.5 a depth average that must survive all of them"""
    data, nr, nb = cb.simulate("tiny", 3.0, seed=17, depth=depth, features_per_mb=200.0)
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); bb.finish()
    m = valid_mask(bb)
    for args in SPARSE_ARGS:
        ref = run_oracle(data, args)
        for batch in ("1", "7"):
            r = run_oracle(data, args, binary=EMU_BIN, kind="emu", env_extra={"CRUMBLE_BATCH_READS": batch})
            assert (r["qual"][m] == ref["qual"][m]).all() and r["bed"] == ref["bed"] and r["counters"] == ref["counters"], (args, batch)


def test_list_free_str_search_equals_find_str():
    """cg_mask_lc_lean (what k_str_items runs: last entry + 16 live slots instead of find_STR's repeat list, scan cut at
    rpos + add + 15) against cg_mask_lc (the full list, pinned to str_finder.c by the golden vectors) on 6e5 random, repeat-rich,
    noisy, N-containing, short and long reads with random CIGARs, positions and -i/-s additions"""
    import subprocess
    from util import ROOT
    r = subprocess.run([str(ROOT / "tests" / "emu" / "str_check"), "200000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches=0" in r.stdout


def test_emulated_pipeline_very_deep_columns():
    """columns deeper than MAX_DEPTH = 20000 (snp_score.c:92, 1493-1500): counted, reported as VDEEP in the BED, not processed - their
    reads keep whatever the other columns decided.  Two 16000x amplicons of 250 bp (26 666 reads pile up on the middle columns)."""
    data, nr, nb = cb.simulate("C4", 0.01, seed=9, amplicon_depth=16000, n_amplicons=2, threads=2)
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); bb.finish()
    m = valid_mask(bb)
    for args in (["-9"], ["-1"]):
        ref = run_oracle(data, args)
        assert ref["bed"].count("VDEEP") >= 50
        r = run_oracle(data, args, binary=EMU_BIN, kind="emu")
        assert (r["qual"][m] == ref["qual"][m]).all() and r["bed"] == ref["bed"] and r["counters"] == ref["counters"]
        r = run_oracle(data, args, binary=EMU_BIN, kind="emu", env_extra={"CRUMBLE_BATCH_READS": "7000"})
        assert (r["qual"][m] == ref["qual"][m]).all() and r["bed"] == ref["bed"] and r["counters"] == ref["counters"]


@pytest.mark.parametrize("name,n_dev,batch", [("tiny", 2, 0), ("tiny", 5, 700), ("tiny", 5, 350), ("tiny", 8, 150), ("c1s", 3, 6000), ("c1s", 8, 0), ("c2s", 4, 9000), ("c4s", 3, 2500)], ids=lambda v: str(v))
def test_multi_device_scheduler_on_emulated_devices(name, n_dev, batch):
    """crumble_b200/csrc/cg_multi.c (the scheduler that spreads one batch over the GPUs of a box) and the CRUMBLE_GPUS path of the host driver,
    with emulated contexts standing in for the devices (EMU_DEVICES): region-shard cuts, read halos, the 128-byte state from shard to shard,
    halo records that turn final in a later shard, event and counter gather - qualities, BED and counters must equal the golden vectors,
    for one call per file (batch 0) and for a chain of calls each of which is spread over the devices."""
    data = sim(name)
    bb = cb.BatchBuilder(pinned=False); bb.add_bam_stream(data); bb.finish()
    m = valid_mask(bb)
    env = {"EMU_DEVICES": "8", "CRUMBLE_GPUS": str(n_dev), "CRUMBLE_BATCH_READS": str(batch if batch else 1 << 30)}
    for args in CHAIN_ARGS:
        exp = GOLD[name]["runs"][args]
        r = run_oracle(data, args.split(), binary=EMU_BIN, kind="emu", env_extra=env)
        assert hashlib.sha256(r["qual"][m].tobytes()).hexdigest() == exp["qual_sha256"], args
        assert r["bed"] == exp["bed"], args
        assert r["counters"] == exp["counters"], args


def test_multi_device_scheduler_edge_cases_on_emulated_devices():
    """odd CIGARs, FUNMAP-placed and unplaced reads, two contigs: cuts between contigs separate independent pieces, cuts inside get a halo"""
    for n_dev in (2, 3, 7):
        for tag in ("l9", "l1B", "l5q30", "l3U35"):
            quals, bed = run_cli_sam(EMU_BIN, EDGE[tag], GDIR / "edge_cases.sam", env_extra={"EMU_DEVICES": "8", "CRUMBLE_GPUS": str(n_dev)})
            exp = [tuple(l.rstrip("\n").split("\t")) for l in open(GDIR / f"edge_cases.{tag}.qual.txt")]
            assert quals == exp, (tag, n_dev)
            assert bed == open(GDIR / f"edge_cases.{tag}.bed").read(), (tag, n_dev)
