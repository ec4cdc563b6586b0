# round 2, run m: ncu --set full of the five heavy kernels at C2 and C3 full size, launch lists of one resident step, e2e timeline
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for w in C2 C3; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^k_column$|^k_cells$|^k_cells_general$|^k_rewrite$|^k_str_items$" -s 15 -c 5 -o gpurun_out/r2m_full_$w -f \
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2m_full_$w.log 2>&1
tail -1 gpurun_out/r2m_full_$w.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2m_launches_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2m_launch_$w.log 2>&1
done
CG_TRACE=1 timeout 300 python bench.py --workload C2 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 2 2> gpurun_out/r2m_trace.err > /dev/null; grep cg_process gpurun_out/r2m_trace.err | tail -26
