# round 1, session 6: full gpu suite + smoke + default bench + reference arm of HEAD
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/m_tests.log 2>&1
tail -5 gpurun_out/m_tests.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/m_smoke.log 2>&1
tail -2 gpurun_out/m_smoke.log
( timeout 600 python bench.py ) > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
cat gpurun_out/m_bench.json
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/m_bench_ref.json 2> gpurun_out/m_bench_ref.err
cat gpurun_out/m_bench_ref.json
