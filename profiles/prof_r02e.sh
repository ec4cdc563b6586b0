# round 2, run e: compact upload planes, N bases in rank space; launch lists for C3/C4 to see where their time goes
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" ) > gpurun_out/r2e_tests.log 2>&1; tail -5 gpurun_out/r2e_tests.log
for w in C2 C4; do
  ( timeout 600 python bench.py --workload $w --no-cpu-baseline --e2e-steps 3 ) > gpurun_out/r2e_bench_$w.json 2> gpurun_out/r2e_bench_$w.err; python - <<PY
import json
try:
    j=json.loads([x for x in open("gpurun_out/r2e_bench_$w.json") if x.startswith("{")][-1])
    print("$w", round(j["ms_per_step"],3), j["config"]["stage_ms"], "e2e_ms", round(j["e2e"]["ms_per_step"],2), "h2d", j["e2e"]["h2d_bytes_per_step"], j["e2e"]["h2d_ms"], j["e2e"]["d2h_ms"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2e_bench_$w.err").read()[-1500:])
PY
done
for w in C3 C4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2e_launch_$w.log 2>&1
python tools/launch_summary.py gpurun_out/r2e_launches_$w.csv 2>/dev/null | head -14
done
CG_TRACE=1 timeout 300 python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/r2e_trace.err > /dev/null; grep cg_process gpurun_out/r2e_trace.err | tail -40
