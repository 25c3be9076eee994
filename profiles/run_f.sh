#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'edge_mean_block' -c 2 \
    -f -o gpurun_out/prof_solve4 python profiles/run_solve.py > gpurun_out/ncu_solve4.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
echo "bench rc=$?" >> gpurun_out/bench_f.err
DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_trace.so TRACE_OBJECTS=36 timeout 300 python profiles/trace_fused.py > gpurun_out/trace.txt 2> gpurun_out/trace.err
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest.txt | tail; tail -2 gpurun_out/bench_f.err
