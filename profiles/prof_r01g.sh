# round 1, session 4: source-level ncu capture of k_column and k_rewrite (C2 x 0.125)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_column|k_rewrite" -s 6 -c 2 -o gpurun_out/prof_g -f python bench.py --workload C2 --scale 0.125 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/g_b.log 2>&1
tail -2 gpurun_out/g_b.log | cut -c1-300
ls -la gpurun_out
