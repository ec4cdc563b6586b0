# round 2, run l: whole GPU suite after the k_epochs search fix + work list for general-CIGAR reads; bench lines
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2l_tests.log
for w in C2 C1 C3 C4; do timeout 600 python bench.py --workload $w > gpurun_out/r2l_bench_$w.json 2> gpurun_out/r2l_bench_$w.err; tail -c 1500 gpurun_out/r2l_bench_$w.json; done
