# round 2, run u (last seconds): the bench line with the parity gate, smallest workload
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 34 python bench.py --workload C4 --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/r2u_bench_C4.json 2> gpurun_out/r2u_bench_C4.err; tail -c 1400 gpurun_out/r2u_bench_C4.json; tail -3 gpurun_out/r2u_bench_C4.err
