#!/bin/bash
# tools/strong.sh N... : on a multi-GPU box — bench.py --region-shards (one contig over N ranks) for each N given
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
for n in "$@"; do
  if [ $n = 1 ]; then timeout 600 python bench.py --region-shards --steps 3 --warmup 3 > gpurun_out/strong_n$n.json 2> gpurun_out/strong_n$n.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --region-shards --steps 3 --warmup 3 > gpurun_out/strong_n$n.json 2> gpurun_out/strong_n$n.err; fi
  echo "== N=$n"; tail -1 gpurun_out/strong_n$n.json | cut -c1-1200; grep -iE "error|Traceback" -A5 gpurun_out/strong_n$n.err | head -20
done
