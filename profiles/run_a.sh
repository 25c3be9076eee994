#!/bin/bash
# first GPU pass of a build: pipe microbenchmark, GPU parity tests, default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 profiles/bin/microbench_pipes > gpurun_out/microbench.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench rc=$?" >> gpurun_out/bench_a.err
tail -5 gpurun_out/pytest.txt; cat gpurun_out/microbench.txt; tail -c 3000 gpurun_out/bench_a.json
