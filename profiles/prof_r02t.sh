# round 2, run t (last seconds): C3 bench line with the event list re-fetched instead of the call re-run
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 45 python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2t_bench_C3.json 2> gpurun_out/r2t_bench_C3.err; tail -c 900 gpurun_out/r2t_bench_C3.json; tail -2 gpurun_out/r2t_bench_C3.err
