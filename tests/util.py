"""Test helpers: run the CPU oracle on a raw BAM stream and bring its output into the
same structure-of-arrays layout the GPU path uses.

Oracle choice: oracle/_ref/crumble_ref (the reference's own sources compiled verbatim,
built where /root/reference exists and shipped as a binary).  A missing binary is an error,
not a downgrade; oracle/bin/crumble_oracle (this repository's plain-C restatement) is used
only where a test asks for it by name."""
import os
import re
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_BIN = ROOT / "oracle" / "_ref" / "crumble_ref"
PORT_BIN = ROOT / "oracle" / "bin" / "crumble_oracle"
EMU_BIN = ROOT / "tests" / "emu" / "emu_crumble"


def oracle_bin(kind=None):
    """kind None = the checker of record: the verbatim reference (oracle/_ref).  There is no silent downgrade to the port: a parity
    claim made against the restatement must say so (kind="port", or CRUMBLE_ORACLE=port in the environment)."""
    if kind is None:
        kind = os.environ.get("CRUMBLE_ORACLE", "reference")
    if kind == "reference":
        if REF_BIN.exists():
            return REF_BIN, "reference"
        raise FileNotFoundError(f"{REF_BIN} is missing: the parity tests compare with the reference's own code. Build it where /root/reference exists "
                                "(make -C oracle ref; it travels to the GPU box with the snapshot) or set CRUMBLE_ORACLE=port to knowingly test against the restatement")
    if kind == "port" and PORT_BIN.exists():
        return PORT_BIN, "port"
    raise FileNotFoundError("no oracle binary: run __graft_entry__.build()")


def parse_counters(stderr_text):
    import crumble_b200 as cb
    pats = {
        "diff": r"A/B Diff\s+= (\d+)", "indel_pair": r"A/B Indel\s+= (\d+) / (\d+)",
        "hetA": r"A:  Het\s+= (\d+) / (\d+)", "homA": r"A:  Hom\s+= (\d+) / (\d+)", "discA": r"A:  Discrep\s+= (\d+)",
        "hetB": r"B:  Het\s+= (\d+) / (\d+)", "homB": r"B:  Hom\s+= (\d+) / (\d+)", "discB": r"B:  Discrep\s+= (\d+)",
        "columns": r"Columns\s+= (\d+)", "low_mqual_perc": r"Low_mqual_perc\s+= (\d+)", "clip_perc": r"Clip_perc\s+= (\d+)",
        "ins_len_perc": r"Ins_len_perc\s+= (\d+)", "indel_ov_perc": r"indel_ov_perc\s+= (\d+)", "over_depth": r"count_over_depth = (\d+)",
    }
    g = {k: re.search(p, stderr_text) for k, p in pats.items()}
    if not all(g.values()):
        return None
    return {
        "diff": int(g["diff"][1]), "indel_qual": int(g["indel_pair"][1]), "indel": int(g["indel_pair"][2]),
        "het_qual_A": int(g["hetA"][1]), "het_A": int(g["hetA"][2]), "hom_qual_A": int(g["homA"][1]), "hom_A": int(g["homA"][2]),
        "discrep_A": int(g["discA"][1]), "het_qual_B": int(g["hetB"][1]), "het_B": int(g["hetB"][2]),
        "hom_qual_B": int(g["homB"][1]), "hom_B": int(g["homB"][2]), "discrep_B": int(g["discB"][1]),
        "columns": int(g["columns"][1]), "low_mqual_perc": int(g["low_mqual_perc"][1]), "clip_perc": int(g["clip_perc"][1]),
        "ins_len_perc": int(g["ins_len_perc"][1]), "indel_ov_perc": int(g["indel_ov_perc"][1]), "over_depth": int(g["over_depth"][1]),
    }


def header_names(data: np.ndarray):
    b = data.tobytes()[: 1 << 20]
    lt = int.from_bytes(b[4:8], "little")
    p = 8 + lt
    n = int.from_bytes(b[p:p + 4], "little"); p += 4
    names = []
    for _ in range(n):
        ln = int.from_bytes(b[p:p + 4], "little"); p += 4
        names.append(b[p:p + ln - 1].decode()); p += ln + 4
    return names


def run_oracle(data: np.ndarray, args, kind=None, binary=None, timing=False, env_extra=None):
    """Run the oracle CLI on an uncompressed BAM stream; returns quals in batch layout, BED text, counters."""
    import crumble_b200 as cb
    if binary is None:
        binary, kind = oracle_bin(kind)
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=tmpdir) as td:
        fin, fout, fbed = os.path.join(td, "in.ubam"), os.path.join(td, "out.ubam"), os.path.join(td, "out.bed")
        data.tofile(fin)
        env = dict(os.environ)
        env["CRUMBLE_REF_TIMING"] = "1"
        env.update(env_extra or {})
        cmd = [str(binary), "-z", "-v"] + list(args) + ["-b", fbed, "-O", "bam,raw", fin, fout]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"oracle failed ({r.returncode}): {r.stderr[-2000:]}")
        out = np.fromfile(fout, dtype=np.uint8)
        bed = open(fbed).read()
    bb = cb.BatchBuilder(pinned=False)
    bb.add_bam_stream(out)
    bb.finish()
    m = re.search(r"transcode_seconds=([0-9.]+)", r.stderr)
    res = {"qual": bb.qual().copy(), "pos": bb.positions(), "len": bb.lengths(), "off": bb.offsets(), "bed": bed, "counters": parse_counters(r.stderr), "names": header_names(data),
           "kind": kind, "seconds": float(m[1]) if m else None, "stderr": r.stderr}
    bb.close()
    return res


def valid_mask(bb):
    """Boolean mask over the padded quality buffer selecting real quality bytes."""
    off, ln = bb.offsets(), bb.lengths()
    total = int(bb.batch.qual_bytes)
    m = np.zeros(total + 1, dtype=np.int32)
    np.add.at(m, off, 1)
    np.add.at(m, off + ln, -1)
    return np.cumsum(m[:-1]) > 0


def add_ambiguity_codes(data: np.ndarray, rate: float, seed: int = 1):
    """a copy of an uncompressed BAM stream with a fraction of the bases replaced by N (mostly), other ambiguity codes and '='"""
    import struct
    d = data.copy(); rng = np.random.default_rng(seed)
    b = d.tobytes(); p = 8 + struct.unpack_from('<i', b, 4)[0]
    nref = struct.unpack_from('<i', b, p)[0]; p += 4
    for _ in range(nref):
        p += 4 + struct.unpack_from('<i', b, p)[0] + 4
    n = 0
    while p + 4 <= len(b):
        bs = struct.unpack_from('<i', b, p)[0]; r = p + 4
        lname = b[r + 8]; ncig = struct.unpack_from('<H', b, r + 12)[0]; lseq = struct.unpack_from('<i', b, r + 16)[0]
        so = r + 32 + lname + 4 * ncig
        for x in rng.integers(0, max(lseq, 1), rng.binomial(lseq, rate) if lseq else 0):
            code = [15, 15, 15, 3, 0, 14][rng.integers(0, 6)]
            i = so + (int(x) >> 1)
            d[i] = (d[i] & 0xf0) | code if x & 1 else (d[i] & 0x0f) | (code << 4)
            n += 1
        p += 4 + bs
    return d, n
