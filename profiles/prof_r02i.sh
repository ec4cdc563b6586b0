# round 2, run i (8 GPUs): strong scaling of C2 over region shards at N=8 and N=4, whole-genome configuration through the C scheduler
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1; nvidia-smi topo -m 2>/dev/null | head -12
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/r2i_n8.json 2> gpurun_out/r2i_n8.err; tail -c 1800 gpurun_out/r2i_n8.json; tail -3 gpurun_out/r2i_n8.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 --no-parity ) > gpurun_out/r2i_n4.json 2> gpurun_out/r2i_n4.err; tail -c 1200 gpurun_out/r2i_n4.json; tail -3 gpurun_out/r2i_n4.err
( timeout 900 python bench.py --workload C5 --gpus 8 --steps 3 --warmup 1 ) > gpurun_out/r2i_c5_n8.json 2> gpurun_out/r2i_c5_n8.err; tail -c 2000 gpurun_out/r2i_c5_n8.json; tail -5 gpurun_out/r2i_c5_n8.err
