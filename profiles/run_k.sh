#!/bin/bash
# full GPU pass: parity suite, the default bench line, launch list + ncu capture of the dominant kernel
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.txt 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/k_pytest.txt; tail -3 gpurun_out/k_pytest.txt
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/k_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r02/launches_r02.log 2>&1
python profiles/summarise.py launches gpurun_out/r02/launches_r02.csv gpurun_out/r02/r02_launches.md
bash profiles/run_j.sh > gpurun_out/k_ncu.log 2>&1
cat gpurun_out/r02/traffic.log; ls gpurun_out/r02
