# round 1, session 5: host timeline of the streamed end-to-end path
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/e2e_trace.py 1.0 > gpurun_out/j_trace.log 2>&1
grep -v "slice enqueued" gpurun_out/j_trace.log | tail -20; grep "slice enqueued" gpurun_out/j_trace.log | tail -4
