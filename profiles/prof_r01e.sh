# round 1, final pass of session 3: tests, bench (both arms), launch list and full ncu captures at the bench's own scale
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/e_tests.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/e_bench.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/e_bench_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/e_b1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_column|k_rewrite|k_flagged" -s 9 -c 3 -o gpurun_out/prof_e_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/e_b2.log 2>&1
tail -3 gpurun_out/e_tests.log; tail -2 gpurun_out/e_bench.log | cut -c1-400; tail -2 gpurun_out/e_bench_ref.log | cut -c1-300
