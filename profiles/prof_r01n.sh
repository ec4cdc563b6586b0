# round 1, session 8 (final): smoke, default bench + reference arm of HEAD, ncu launch list of the bench command,
# one ncu --set full capture (with source) of the three heavy kernels at the bench's size class
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/n_smoke.log 2>&1
tail -2 gpurun_out/n_smoke.log
( timeout 900 python bench.py ) > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
cat gpurun_out/n_bench.json
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/n_bench_ref.json 2> gpurun_out/n_bench_ref.err
cat gpurun_out/n_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/n_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/n_launch_bench.log 2>&1
tail -1 gpurun_out/n_launch_bench.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_column|k_rewrite|k_str_items" -s 9 -c 3 -o gpurun_out/n_full -f \
    python bench.py --workload C2 --scale 0.125 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/n_full.log 2>&1
tail -1 gpurun_out/n_full.log | cut -c1-120
