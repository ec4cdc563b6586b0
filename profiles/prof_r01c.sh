cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_column -s 3 -c 1 -o gpurun_out/prof_column_c -f python bench.py --workload C2 --scale 0.125 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b2.log 2>&1
