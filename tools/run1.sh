cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/t_tests.log 2>&1; tail -15 gpurun_out/t_tests.log
CG_TRACE=1 timeout 600 python bench.py --workload C2 --scale 0.25 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/t_bench.log 2>&1; grep cg_process gpurun_out/t_bench.log | tail -24; tail -1 gpurun_out/t_bench.log | cut -c1-300
