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
