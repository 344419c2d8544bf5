cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_tc_2gpu.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
