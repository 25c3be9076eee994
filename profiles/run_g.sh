#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
echo "bench rc=$?" >> gpurun_out/bench_g.err
timeout 900 python bench.py --config sweep1m --steps 2 --warmup 3 --graph-collective > gpurun_out/bench_sweep1m_1gpu.json 2> gpurun_out/bench_sweep1m.err
timeout 600 python bench.py --config kitti_val_full --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_kitti_val_full_1gpu.json 2> gpurun_out/bench_full.err
DCD_B200_KTC=1 timeout 300 python -m pytest tests/test_gpu_gmw.py tests/test_gpu_transport.py -m gpu -q -k "transport or edge_p or training_step" > gpurun_out/pytest_ktc.txt 2>&1
bash profiles/run_profile_r02.sh > gpurun_out/profile.log 2>&1
tail -2 gpurun_out/bench_g.err; tail -4 gpurun_out/pytest_ktc.txt; tail -12 gpurun_out/profile.log
