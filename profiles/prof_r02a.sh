# round 2, run a: baseline of HEAD before the kernel work — full GPU test suite (with the new whole-file C2/C3 reference parity),
# bench lines for C1..C4, the same-config reference arm, and ncu --set full of the three heavy kernels at C2 and C3 full size
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2a_tests.log 2>&1; tail -14 gpurun_out/r2a_tests.log
for w in C2 C1 C3 C4; do
  ( timeout 600 python bench.py --workload $w ) > gpurun_out/r2a_bench_$w.json 2> gpurun_out/r2a_bench_$w.err; cut -c1-1500 gpurun_out/r2a_bench_$w.json
done
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2a_ref_C2.json 2> gpurun_out/r2a_ref_C2.err; cut -c1-600 gpurun_out/r2a_ref_C2.json
for w in C2 C3; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_column|k_rewrite|k_str_items" -s 9 -c 3 -o gpurun_out/r2a_full_$w -f \
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2a_full_$w.log 2>&1
tail -1 gpurun_out/r2a_full_$w.log | cut -c1-200
done
