"""BASELINE.json's full-size configuration (synthetic chr20, 64 Mb at 30x: 1.29e7 reads, 1.92e9 aligned bases) on the GPU.
The CPU oracle needs minutes for the whole of it, so parity at this size is checked through properties:
  * slicing invariance: resident (upload + run + download, one slice) and streamed (cg_process, two chunk sizes, up to 32
    slices with the cross-slice state carried on the device) give the same qualities, BED events and counters;
  * windowed oracle parity: for three 150 kb windows the verbatim reference is run on the window alone (-r) and every read
    lying at least 5 kb inside it must carry exactly the qualities the full-size GPU run gave it;
  * counter sanity: the number of columns the path counts equals the number of covered reference positions."""
import numpy as np
import pytest

import crumble_b200 as cb
from util import header_names, run_oracle, valid_mask

pytestmark = pytest.mark.gpu


def _bytes_of(off, ln):
    """indices of the quality bytes of the given records (offset, length)"""
    tot = int(ln.sum())
    starts = np.repeat(off - np.concatenate(([0], np.cumsum(ln)[:-1])), ln)
    return starts + np.arange(tot, dtype=np.int64)


def test_full_size_c2_properties():
    data, n_reads, n_bases = cb.simulate("C2", 1.0, seed=2)
    assert n_reads > 12_000_000 and n_bases > 1_900_000_000
    bb = cb.BatchBuilder()
    bb.add_bam_stream(data)
    batch = bb.finish()
    mask = valid_mask(bb)
    g = cb.Crumble(cb.default_params(9), device=0)

    # --- slicing invariance ---
    g.upload(batch); g.run()
    res = g.download(batch)
    ncols = g.n_columns()
    outs = []
    for chunk in (96 << 20, 40 << 20):
        g.set_chunk_bytes(chunk)
        o = g.process(batch)
        assert np.array_equal(o["qual"][mask], res["qual"][mask]), f"streamed ({chunk >> 20} MiB chunks) differs from resident"
        assert o["counters"] == res["counters"]
        outs.append(o)
    assert np.array_equal(outs[0]["events"], outs[1]["events"]) and len(outs[0]["events"]) == res["n_events"]
    assert res["counters"]["columns"] == ncols                      # every covered position is a counted column at -9
    full = outs[0]["qual"].copy()
    del outs, res
    pos, ln, off = bb.positions(), bb.lengths().astype(np.int64), bb.offsets()

    # --- windowed oracle parity ---
    name = header_names(data)[0]
    margin, width = 5000, 150_000
    checked = 0
    for beg in (1_000_000, 9_000_000, 20_000_000):
        end = beg + width
        ref = run_oracle(data, ["-9", "-r", f"{name}:{beg + 1}-{end}"])
        # records starting inside the window appear in both, in the same order
        i0, i1 = int(np.searchsorted(pos, beg, side="left")), int(np.searchsorted(pos, end, side="left"))
        j0, j1 = int(np.searchsorted(ref["pos"], beg, side="left")), int(np.searchsorted(ref["pos"], end, side="left"))
        assert i1 - i0 == j1 - j0 and i1 - i0 > 20_000
        assert np.array_equal(pos[i0:i1], ref["pos"][j0:j1]) and np.array_equal(ln[i0:i1], ref["len"][j0:j1])
        inside = (pos[i0:i1] >= beg + margin) & (pos[i0:i1] + 1000 <= end - margin)
        a = full[_bytes_of(off[i0:i1][inside], ln[i0:i1][inside])]
        b = ref["qual"][_bytes_of(ref["off"][j0:j1][inside], ref["len"][j0:j1][inside].astype(np.int64))]
        nbad = int((a != b).sum())
        assert nbad == 0, f"window {beg}-{end}: {nbad} quality bytes differ from the reference ({ref['kind']}) run on the window"
        checked += int(inside.sum())
    assert checked > 60_000
    g.close(); bb.close()


@pytest.mark.parametrize("preset,seed,args", [("C1", 1, ["-9"]), ("C4", 4, ["-9"]), ("C4", 4, ["-1"]), ("C1", 1, ["-1", "-B"])],
                         ids=lambda v: "".join(v) if isinstance(v, list) else str(v))
def test_full_size_c1_c4_exact(preset, seed, args):
    """BASELINE.json configs[0] (1 Mb at 30x) and configs[3] (1000x amplicon panel: hot columns, STR-rich, every read of an amplicon
    starting at the same column) at their full sizes, where the CPU oracle still finishes in seconds: bit-exact qualities, BED
    lines and counters, resident and streamed."""
    from test_gpu_parity import params_from_args
    data, n_reads, n_bases = cb.simulate(preset, 1.0, seed=seed)
    bb = cb.BatchBuilder(); bb.add_bam_stream(data); batch = bb.finish()
    mask = valid_mask(bb)
    ref = run_oracle(data, args)
    g = cb.Crumble(params_from_args(args), device=0)
    g.set_chunk_bytes(8 << 20)
    out = g.process(batch)
    nbad = int((out["qual"][mask] != ref["qual"][mask]).sum())
    assert nbad == 0, f"{nbad} quality bytes differ from the oracle ({ref['kind']})"
    assert cb.bed_text(out["events"], ref["names"]) == ref["bed"]
    assert out["counters"] == ref["counters"]
    g.upload(batch); g.run(); res = g.download(batch)
    assert np.array_equal(res["qual"][mask], out["qual"][mask]) and res["counters"] == out["counters"]
    g.close(); bb.close()
