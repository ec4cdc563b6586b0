# round 2, run d: list-free STR search (k_flagged emits an item list, k_str_items one thread per item), full GPU suite, bench C2/C3/C4
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_tests.log 2>&1; tail -5 gpurun_out/r2d_tests.log
for w in C2 C3 C4 C1; do
  ( timeout 600 python bench.py --workload $w --no-cpu-baseline --e2e-steps 2 ) > gpurun_out/r2d_bench_$w.json 2> gpurun_out/r2d_bench_$w.err; python - <<PY
import json
try:
    j=json.loads([x for x in open("gpurun_out/r2d_bench_$w.json") if x.startswith("{")][-1])
    print("$w", round(j["ms_per_step"],3), j["config"]["stage_ms"], "e2e_ms", round(j["e2e"]["ms_per_step"],2))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2d_bench_$w.err").read()[-1500:])
PY
done
