# round 2, run g: shard phases + multi-device scheduler (contexts on one GPU), unroll 1
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" ) > gpurun_out/r2g_tests.log 2>&1; tail -25 gpurun_out/r2g_tests.log
( timeout 600 python bench.py --workload C2 --no-cpu-baseline --e2e-steps 3 ) > gpurun_out/r2g_bench_C2.json 2> gpurun_out/r2g_bench_C2.err; python - <<PY
import json
try:
    j=json.loads([x for x in open("gpurun_out/r2g_bench_C2.json") if x.startswith("{")][-1])
    print("C2", round(j["ms_per_step"],3), j["config"]["stage_ms"], "e2e_ms", round(j["e2e"]["ms_per_step"],2))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2g_bench_C2.err").read()[-1500:])
PY
