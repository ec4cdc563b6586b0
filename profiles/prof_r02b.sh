# round 2, run b: new column stage (k_cells + k_column with cp.async staging, dynamic tile scheduling, lean finalisation)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r2b_smoke.log 2>&1; tail -3 gpurun_out/r2b_smoke.log
( timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_size" ) > gpurun_out/r2b_tests.log 2>&1; tail -8 gpurun_out/r2b_tests.log
for w in C2 C4 C3; do
  ( timeout 600 python bench.py --workload $w --no-cpu-baseline ) > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err; cut -c1-1200 gpurun_out/r2b_bench_$w.json; tail -3 gpurun_out/r2b_bench_$w.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_column|k_cells" -s 9 -c 3 -o gpurun_out/r2b_full_C2 -f \
    python bench.py --workload C2 --scale 0.25 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2b_full_C2.log 2>&1
tail -1 gpurun_out/r2b_full_C2.log | cut -c1-200
