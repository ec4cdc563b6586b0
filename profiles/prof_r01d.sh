# round-1 session-3 GPU pass: tests, bench, launch list, full ncu capture of the two top kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/d_smi.log 2>&1
nproc >> gpurun_out/d_smi.log; lscpu | grep "Model name" >> gpurun_out/d_smi.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/d_tests.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/d_bench.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/d_bench_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d_launches.csv python bench.py --workload C2 --scale 0.125 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/d_b1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_column|k_rewrite" -s 6 -c 2 -o gpurun_out/prof_d -f python bench.py --workload C2 --scale 0.125 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/d_b2.log 2>&1
tail -3 gpurun_out/d_tests.log; tail -2 gpurun_out/d_bench.log; tail -2 gpurun_out/d_bench_ref.log
