#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
echo "bench rc=$?" >> gpurun_out/bench_g.err
timeout 900 python bench.py --config sweep1m --steps 2 --warmup 3 --graph-collective > gpurun_out/bench_sweep1m_1gpu.json 2> gpurun_out/bench_sweep1m.err
timeout 600 python bench.py --config kitti_val_full --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_kitti_val_full_1gpu.json 2> gpurun_out/bench_full.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
bash profiles/run_profile_r02.sh > gpurun_out/profile.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest.txt | tail; tail -2 gpurun_out/bench_g.err; tail -2 gpurun_out/smoke.txt; tail -12 gpurun_out/profile.log
