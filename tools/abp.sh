#!/bin/bash
# tools/abp.sh [KERNEL_REGEX] : on the GPU box — parity tests, stage timings at C2 x SCALE (default 1), then one source-level
# ncu capture (C2 x 0.125) of the kernels matching KERNEL_REGEX (default k_column|k_rewrite) -> gpurun_out/prof_p.ncu-rep
cd ${GRAFT_REPO_ROOT:-.}
SCALE=${SCALE:-1} bash tools/ab.sh
[ "$1" = none ] && exit 0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${1:-k_column|k_rewrite}" -s ${SKIP:-6} -c ${COUNT:-2} -o gpurun_out/prof_p -f python bench.py --workload C2 --scale 0.125 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/p_b.log 2>&1
tail -1 gpurun_out/p_b.log | cut -c1-200
