# round 2, run q (last GPU minutes of the round): compact planes of the per-record arrays on the device: the tests that use packed batches, one C2 bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k "compact_planes or refetched or ambiguity" > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2q_tests.log
timeout 120 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2q_bench_C2.json 2> gpurun_out/r2q_bench_C2.err; python - <<PY
import json
try:
    j=json.loads([x for x in open("gpurun_out/r2q_bench_C2.json") if x.startswith("{")][-1])
    print("C2", round(j["ms_per_step"],3), j["config"]["stage_ms"], "e2e_ms", round(j["e2e"]["ms_per_step"],2), "h2d", j["e2e"]["h2d_bytes_per_step"], j["e2e"]["h2d_ms"], j["e2e"]["d2h_ms"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2q_bench_C2.err").read()[-1500:])
PY
