#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_full.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_full.txt
grep -E "passed|failed|^FAILED|^ERROR|ours vs f64|worst|cg iterations|step [01] loss|a~c|20x" gpurun_out/pytest_full.txt | grep -v "print(" > gpurun_out/pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'edge_mean_block' -c 2 \
    -f -o gpurun_out/prof_solve3 python profiles/run_solve.py > gpurun_out/ncu_solve3.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
echo "bench rc=$?" >> gpurun_out/bench_e.err
cat gpurun_out/pytest.txt | tail -40
tail -2 gpurun_out/bench_e.err
