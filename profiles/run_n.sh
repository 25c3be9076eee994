#!/bin/bash
# 2-GPU check of the final build: NCCL parity tests + the default bench line as the driver launches it at N = 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/n_pytest_dist.txt 2>&1; tail -3 gpurun_out/n_pytest_dist.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_kitti_val_2gpu.json 2> gpurun_out/bench_kitti_val_2gpu.err
echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_kitti_val_2gpu.json; grep -o '"shard_check": {[^}]*}' gpurun_out/bench_kitti_val_2gpu.json; tail -2 gpurun_out/bench_kitti_val_2gpu.err
