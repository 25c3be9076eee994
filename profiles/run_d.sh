#!/bin/bash
# GPU pass: parity tests, ncu of the blocked edge-mean kernel + select, default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'edge_mean_block|edge_select_radix' -c 6 \
    -f -o gpurun_out/prof_solve2 python profiles/run_solve.py > gpurun_out/ncu_solve2.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
echo "bench rc=$?" >> gpurun_out/bench_d.err
grep -E "passed|failed" gpurun_out/pytest.txt | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest.txt | head -20
tail -2 gpurun_out/ncu_solve2.log; tail -2 gpurun_out/bench_d.err
