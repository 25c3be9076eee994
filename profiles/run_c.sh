#!/bin/bash
# GPU pass: parity tests, then the configs[3] / configs[4] bench lines on one GPU
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.txt
timeout 900 python bench.py --config stress256 --steps 2 --warmup 3 > gpurun_out/bench_stress256_1gpu.json 2> gpurun_out/bench_stress256.err
echo "stress rc=$?" >> gpurun_out/bench_stress256.err
timeout 1500 python bench.py --config sweep1m --steps 2 --warmup 3 > gpurun_out/bench_sweep1m_1gpu.json 2> gpurun_out/bench_sweep1m.err
echo "sweep rc=$?" >> gpurun_out/bench_sweep1m.err
grep -E "passed|failed" gpurun_out/pytest.txt | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest.txt | head -20
tail -3 gpurun_out/bench_stress256.err; tail -3 gpurun_out/bench_sweep1m.err
