# round 1, session 5: e2e after moving the per-slice count readback off the D2H copy engine (mapped pinned memory) + tapered last chunks
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/k_tests.log 2>&1
tail -3 gpurun_out/k_tests.log
timeout 600 python tools/e2e_trace.py 1.0 > gpurun_out/k_trace.log 2>&1
grep -v "slice enqueued" gpurun_out/k_trace.log | tail -12; grep "slice enqueued" gpurun_out/k_trace.log | tail -4
