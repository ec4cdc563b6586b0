# round 2, run n: bit-parallel STR search (cg_mask_lc_bits) in k_str_items: GPU suite, bench C2/C3/C4, ncu of k_str_items at C3
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2n_tests.log
for w in C2 C3 C4; do timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2n_bench_$w.json 2> gpurun_out/r2n_bench_$w.err; python - <<PY
import json
try:
    j=json.loads([x for x in open("gpurun_out/r2n_bench_$w.json") if x.startswith("{")][-1])
    print("$w", round(j["ms_per_step"],3), j["config"]["stage_ms"], "e2e_ms", round(j["e2e"]["ms_per_step"],2))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2n_bench_$w.err").read()[-1500:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_str_items$" -s 3 -c 1 -o gpurun_out/r2n_full_C3 -f \
    python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2n_full_C3.log 2>&1
