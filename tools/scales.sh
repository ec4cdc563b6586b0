#!/bin/bash
# tools/scales.sh : stage timings of the default build at C2 x 0.125 .. 1 (does the per-unit cost depend on the batch size?)
cd ${GRAFT_REPO_ROOT:-.}
for s in 0.125 0.25 0.5 1; do
  timeout 600 python bench.py --workload C2 --scale $s --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/sc_$s.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/sc_$s.log") if x.startswith("{")][-1]; j=json.loads(l)
print("$s", j["ms_per_step"], j["config"]["stage_ms"])
PY
done
