# round 2, run r (2 GPUs, last minutes): the shard bench with the ring mailbox and device-clock timing at N=2
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( CG_BENCH_PHASES=1 timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity ) > gpurun_out/r2r_n2.json 2> gpurun_out/r2r_n2.err; tail -c 1500 gpurun_out/r2r_n2.json; grep phases gpurun_out/r2r_n2.err; tail -2 gpurun_out/r2r_n2.err | cut -c1-300
