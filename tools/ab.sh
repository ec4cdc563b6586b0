#!/bin/bash
# tools/ab.sh [variant ...] : on the GPU box — parity tests on the default build, then stage timings of the default
# build and of each named variant (crumble_b200/lib/variants/libcrumble_gpu_<v>.so) on C2 x SCALE.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
SCALE=${SCALE:-0.25}
( timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} ) > gpurun_out/ab_tests.log 2>&1
tail -4 gpurun_out/ab_tests.log
for v in default "$@"; do
  if [ $v = default ]; then unset CRUMBLE_GPU_LIB; else export CRUMBLE_GPU_LIB=$PWD/crumble_b200/lib/variants/libcrumble_gpu_$v.so; fi
  timeout 600 python bench.py --workload ${WORKLOAD:-C2} --scale $SCALE --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ab_$v.log 2>&1
  echo "== $v"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/ab_$v.log") if x.startswith("{")][-1]; j=json.loads(l)
    print(j["ms_per_step"], j["config"]["stage_ms"], "e2e_ms", j["e2e"]["ms_per_step"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/ab_$v.log").read()[-1500:])
PY
done
