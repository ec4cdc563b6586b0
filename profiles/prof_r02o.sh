# round 2, run o: k_rewrite with direct global loads of qualities / sequences (RW_DIRECT=1, default) against the bulk-staged form (rw0);
# CgRead 16-byte aligned; e2e timeline of C3 (why is it 114 ms when the two copy spans are 53 and 50 ms?)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
PYTEST_K="not full_size" SCALE=1.0 bash tools/ab.sh rw0
CG_TRACE=1 timeout 300 python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/r2o_trace_C3.err > /dev/null; grep cg_process gpurun_out/r2o_trace_C3.err | tail -30
