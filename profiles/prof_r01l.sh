# round 1, session 5: full gpu suite incl. the -S/-k/-K/-N/-y/-R option tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/l_tests.log 2>&1
tail -15 gpurun_out/l_tests.log
