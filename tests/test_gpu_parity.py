"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Bar: quality strings, BED events and -v counters bit-exact."""
import numpy as np
import pytest

import crumble_b200 as cb
from util import run_oracle, valid_mask

pytestmark = pytest.mark.gpu

LEVELS = [["-9"], ["-8"], ["-7"], ["-5"], ["-3"], ["-1"], ["-1", "-B"], ["-9", "-Y0.1"], ["-5", "-q30"],
          ["-9", "-U30"], ["-1", "-m10", "-C0.05", "-Z0.01"]]


def params_from_args(args):
    """Python mirror of the reference getopt loop for the options the tests use."""
    p = cb.default_params()
    lib = cb.load_lib()
    import ctypes as C
    for a in args:
        k, v = a[1], a[2:]
        if k in "135789":
            lib.cg_params_level(C.byref(p), int(k))
        elif k == "B": p.binary_qual = 1
        elif k == "Y": p.indel_fract = float(v)
        elif k == "q": p.min_qual_A = int(v)
        elif k == "Q": p.min_qual_B = int(v)
        elif k == "U": p.qcap = int(v)
        elif k == "m": p.min_mqual = int(v)
        elif k == "C": p.clip_perc = float(v)
        elif k == "Z": p.ins_len_perc = float(v)
        elif k == "P": p.over_depth = float(v)
        elif k == "p": p.pblock = int(v)
        elif k == "L": p.reduce_qual = int(v)
        else: raise ValueError(a)
    return p


_cache = {}


def dataset(name):
    if name not in _cache:
        preset, scale, seed = {"tiny": ("tiny", 1.0, 3), "c1s": ("C1", 0.25, 11), "c2s": ("C2", 1 / 128, 5), "c4s": ("C4", 0.05, 4)}[name]
        data, nr, nb = cb.simulate(preset, scale, seed)
        bb = cb.BatchBuilder()
        bb.add_bam_stream(data)
        batch = bb.finish()
        _cache[name] = (data, bb, batch, valid_mask(bb))
    return _cache[name]


def check(name, args):
    data, bb, batch, mask = dataset(name)
    g = cb.Crumble(params_from_args(args), device=0)
    out = g.process(batch)
    ref = run_oracle(data, args)
    nbad = int((out["qual"][mask] != ref["qual"][mask]).sum())
    assert nbad == 0, f"{nbad} quality bytes differ from the oracle ({ref['kind']})"
    assert cb.bed_text(out["events"], ref["names"]) == ref["bed"]
    assert out["counters"] == ref["counters"]
    g.close()


@pytest.mark.parametrize("args", LEVELS, ids=lambda a: "".join(a))
def test_tiny_all_levels(args):
    check("tiny", args)


@pytest.mark.parametrize("name", ["c1s", "c2s", "c4s"])
@pytest.mark.parametrize("args", [["-9"], ["-1", "-B"], ["-5"]], ids=lambda a: "".join(a))
def test_configs(name, args):
    check(name, args)


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
