# round 2, run h (2 GPUs): strong-scaling bench line over region shards, small then full size; reference arm untouched
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | head -4; nproc; free -g | head -2
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --scale 0.125 ) > gpurun_out/r2h_n2_small.json 2> gpurun_out/r2h_n2_small.err; tail -c 2500 gpurun_out/r2h_n2_small.json; tail -5 gpurun_out/r2h_n2_small.err
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2h_n2.json 2> gpurun_out/r2h_n2.err; tail -c 2500 gpurun_out/r2h_n2.json; tail -5 gpurun_out/r2h_n2.err
( timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2h_n1.json 2> gpurun_out/r2h_n1.err; tail -c 1500 gpurun_out/r2h_n1.json
