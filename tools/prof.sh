#!/bin/bash
# tools/prof.sh KERNEL_REGEX NAME : one ncu --set full capture of the named kernel(s) (4th bench step) -> gpurun_out/NAME.ncu-rep
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${SKIP:-3} -c ${COUNT:-1} -o gpurun_out/$2 -f python bench.py --workload ${WORKLOAD:-C2} --scale ${SCALE:-0.125} --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log | cut -c1-200
