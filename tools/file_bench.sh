#!/bin/bash
# tools/file_bench.sh [SCALE] : on the GPU box — the crumble_gpu command line on a C2 x SCALE file, BAM in / BAM out:
# raw (uncompressed) and BGZF, one call vs chained calls; prints transcode_gpu's own phase summary and reads/s.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
SCALE=${1:-0.25}
D=/dev/shm/fb; mkdir -p $D
python - <<PY
import sys, time
sys.path.insert(0, ".")
import crumble_b200 as cb
t = time.time()
data, nr, nb = cb.simulate("C2", $SCALE, seed=7)
data.tofile("$D/in.ubam")
print("generated", nr, "reads", nb, "bases", data.nbytes >> 20, "MiB raw BAM in %.1f s" % (time.time() - t))
open("$D/n", "w").write("%d %d" % (nr, nb))
PY
CLI=crumble_b200/lib/crumble_gpu
run() { # label, env, args...
  local label=$1; shift; local envs=$1; shift
  local t0=$(date +%s.%N)
  env CRUMBLE_TIMING=1 $envs $CLI -z -9 "$@" 2> $D/err.txt; local rc=$?
  local t1=$(date +%s.%N)
  python - <<PY
nr, nb = map(int, open("$D/n").read().split())
dt = $t1 - $t0
print("== $label rc=$rc wall %.2f s  %.2f Mreads/s  %.1f Mbases/s" % (dt, nr / dt / 1e6, nb / dt / 1e6))
print(open("$D/err.txt").read().strip()[-700:])
PY
}
run "warm-up (raw->raw, one call)" "CRUMBLE_BATCH_READS=100000000" -O bam,raw $D/in.ubam $D/o1.ubam
run "raw->raw, one call" "CRUMBLE_BATCH_READS=100000000" -O bam,raw $D/in.ubam $D/o1.ubam
run "raw->raw, chained (default 512Ki)" "" -O bam,raw $D/in.ubam $D/o2.ubam
run "raw->raw, chained 128Ki" "CRUMBLE_BATCH_READS=131072" -O bam,raw $D/in.ubam $D/o3.ubam
cmp $D/o1.ubam $D/o2.ubam && cmp $D/o1.ubam $D/o3.ubam && echo "chained output identical to the single call"
run "raw->bgzf (16 threads)" "" -O bam $D/in.ubam $D/o4.bam
run "bgzf->bgzf (16 threads)" "" -O bam $D/o4.bam $D/o5.bam
run "bgzf->bgzf (1 thread)" "HTS_LITE_THREADS=1" -O bam $D/o4.bam $D/o6.bam
cmp $D/o5.bam $D/o6.bam && echo "bgzf output independent of thread count"
nproc; ls -la $D | head -12
rm -rf $D
