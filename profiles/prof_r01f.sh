# round 1, session 4 entry check: GPU parity tests + default bench on the restored tree
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_tests.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/f_bench.log 2>&1
tail -3 gpurun_out/f_tests.log; tail -2 gpurun_out/f_bench.log | cut -c1-1500
