# round 2, run s (4 GPUs, last minutes): the shard bench at N=4
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( CG_BENCH_PHASES=1 timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 --no-parity ) > gpurun_out/r2s_n4.json 2> gpurun_out/r2s_n4.err; tail -c 600 gpurun_out/r2s_n4.json; grep phases gpurun_out/r2s_n4.err | sort -u | cut -c1-120; tail -2 gpurun_out/r2s_n4.err | cut -c1-300
