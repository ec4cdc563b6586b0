# round 1, session 4: 2-GPU scaling sanity (same launch line as the driver's)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/h_bench2.log 2>&1
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/h_bench2_ref.log 2>&1
grep '^{' gpurun_out/h_bench2.log | cut -c1-600; grep '^{' gpurun_out/h_bench2_ref.log | cut -c1-300; tail -5 gpurun_out/h_bench2.log | cut -c1-300
