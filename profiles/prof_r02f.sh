# round 2, run f: k_column after moving the N path out of line; unroll A/B
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
PYTEST_K="ambiguity or compact or tiny_all" SCALE=1.0 bash tools/ab.sh unroll1 unroll4
