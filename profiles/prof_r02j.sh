# round 2, run j: compact download (mask + exceptions, expanded by host threads), VDEEP / aux-tag tests
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" ) > gpurun_out/r2j_tests.log 2>&1; tail -25 gpurun_out/r2j_tests.log
for v in "" "CG_FLAT_D2H=1"; do
( env $v timeout 600 python bench.py --workload C2 --no-cpu-baseline --e2e-steps 4 --steps 5 ) > gpurun_out/r2j_bench_C2.json 2> gpurun_out/r2j_bench_C2.err; python - <<PY
import json
try:
    j=json.loads([x for x in open("gpurun_out/r2j_bench_C2.json") if x.startswith("{")][-1])
    print("C2 [$v]", round(j["ms_per_step"],3), j["config"]["stage_ms"], "e2e_ms", round(j["e2e"]["ms_per_step"],2), "h2d", j["e2e"]["h2d_bytes_per_step"], round(j["e2e"]["h2d_ms"],1), round(j["e2e"]["d2h_ms"],1))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2j_bench_C2.err").read()[-2500:])
PY
done
CG_TRACE=1 timeout 300 python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/r2j_trace.err > /dev/null; grep cg_process gpurun_out/r2j_trace.err | tail -26
