# round 2, run c: tuned cells / column kernels (branch-free fast row, exp for y <= 0, claims three tiles ahead), occupancy A/B, ncu
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
PYTEST_K="not full_size and not chained" SCALE=1.0 bash tools/ab.sh minb4 minb6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_column|k_cells" -s 9 -c 3 -o gpurun_out/r2c_full_C2 -f \
    python bench.py --workload C2 --scale 0.25 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2c_full_C2.log 2>&1
tail -1 gpurun_out/r2c_full_C2.log | cut -c1-200
