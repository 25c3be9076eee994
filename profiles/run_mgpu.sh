#!/bin/bash
# usage: run_mgpu.sh N  — multi-GPU bench lines of BASELINE configs[3] (strong scaling) and configs[4] on N GPUs of one box
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "4" ]; then
  timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_dist.txt 2>&1; tail -3 gpurun_out/pytest_dist.txt
fi
timeout 900 $TR bench.py --gpus $N --config sweep1m --steps 2 --warmup 3 --quick > gpurun_out/bench_sweep1m_${N}gpu.json 2> gpurun_out/bench_sweep1m_${N}gpu.err
echo "sweep rc=$?"
timeout 600 $TR bench.py --gpus $N --config stress256 --steps 2 --warmup 3 > gpurun_out/bench_stress256_${N}gpu.json 2> gpurun_out/bench_stress256_${N}gpu.err
echo "stress rc=$?"
tail -c 600 gpurun_out/bench_sweep1m_${N}gpu.json; echo; tail -c 300 gpurun_out/bench_stress256_${N}gpu.json; echo
tail -3 gpurun_out/bench_sweep1m_${N}gpu.err
