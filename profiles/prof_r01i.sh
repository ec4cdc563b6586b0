# round 1, session 5: state check of HEAD — gpu parity tests, full bench (C2), reference arm, ncu launch list of the bench command
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/i_tests.log 2>&1
tail -3 gpurun_out/i_tests.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/i_bench.log 2>&1
grep '^{' gpurun_out/i_bench.log | cut -c1-2500
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/i_bench_ref.log 2>&1
grep '^{' gpurun_out/i_bench_ref.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/i_ncu_bench.log 2>&1
tail -2 gpurun_out/i_ncu_bench.log | cut -c1-300
