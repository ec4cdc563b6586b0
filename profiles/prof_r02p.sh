# round 2, run p (8 GPUs): strong scaling of C2 over region shards at N=8 with the per-phase host timeline, whole-genome configuration (C5) through the C scheduler
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
( CG_BENCH_PHASES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r2p_n8.json 2> gpurun_out/r2p_n8.err; tail -c 1800 gpurun_out/r2p_n8.json; grep phases gpurun_out/r2p_n8.err; tail -3 gpurun_out/r2p_n8.err
( timeout 900 python bench.py --workload C5 --gpus 8 --steps 3 --warmup 1 ) > gpurun_out/r2p_c5_n8.json 2> gpurun_out/r2p_c5_n8.err; tail -c 2000 gpurun_out/r2p_c5_n8.json; tail -5 gpurun_out/r2p_c5_n8.err
