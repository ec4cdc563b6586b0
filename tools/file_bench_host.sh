#!/bin/bash
# tools/file_bench_host.sh [SCALE] : the HOST side of the command line alone (reader, batcher, finalise, writer), no GPU needed:
# tests/emu/emu_crumble with EMU_NULL=1 hands the qualities back untouched, so what is timed is record decode, batching, copies and
# encode.  Prints transcode_gpu's own phase summary for raw and BGZF input / output.
cd "$(dirname "$0")/.."
SCALE=${1:-0.0625}
D=/dev/shm/fbh; mkdir -p $D
make -s -C tests/emu emu_crumble
python - <<PY
import sys
sys.path.insert(0, ".")
import crumble_b200 as cb
data, nr, nb = cb.simulate("C2", $SCALE, seed=7)
data.tofile("$D/in.ubam"); open("$D/n", "w").write("%d %d" % (nr, nb))
PY
CLI=tests/emu/emu_crumble
run() { local label=$1; shift
  local t0=$(date +%s.%N); EMU_NULL=1 CRUMBLE_TIMING=1 $CLI -z -9 "$@" 2> $D/err.txt; local t1=$(date +%s.%N)
  python - <<PY
nr, nb = map(int, open("$D/n").read().split()); dt = $t1 - $t0
print("== $label wall %.2f s  %.2f Mreads/s  %.1f Mbases/s" % (dt, nr / dt / 1e6, nb / dt / 1e6)); print(open("$D/err.txt").read().strip()[-400:])
PY
}
run "raw->raw" -O bam,raw $D/in.ubam $D/o1.ubam
run "raw->bgzf" -O bam $D/in.ubam $D/o2.bam
run "bgzf->bgzf" -O bam $D/o2.bam $D/o3.bam
nproc; rm -rf $D
