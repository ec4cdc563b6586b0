cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
CG_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 --e2e-steps 3 > gpurun_out/t_bench_full.log 2>&1; grep cg_process gpurun_out/t_bench_full.log | tail -26; tail -1 gpurun_out/t_bench_full.log
