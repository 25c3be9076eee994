#!/bin/bash
# 8-GPU records of the final build: configs[3] (strong scaling, all-gather inside the timed region) and the default line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --config sweep1m --steps 2 --warmup 3 --quick > gpurun_out/bench_sweep1m_8gpu.json 2> gpurun_out/bench_sweep1m_8gpu.err
echo "sweep rc=$?"; grep "^{" gpurun_out/bench_sweep1m_8gpu.json | cut -c1-200
timeout 200 $TR bench.py --gpus 8 --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_kitti_val_8gpu.json 2> gpurun_out/bench_kitti_val_8gpu.err
echo "kitti rc=$?"; grep "^{" gpurun_out/bench_kitti_val_8gpu.json | cut -c1-200
