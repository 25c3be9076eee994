#!/bin/bash
# GPU pass: parity tests, ncu capture of the DGDE-side kernels, in-kernel timeline of the fused forward, default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'edge_mean_group|edge_select_radix|edge_solve_bwd' -c 8 \
    -f -o gpurun_out/prof_solve python profiles/run_solve.py > gpurun_out/ncu_solve.log 2>&1
DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_trace.so TRACE_OBJECTS=36 timeout 300 python profiles/trace_fused.py > gpurun_out/trace.txt 2> gpurun_out/trace.err
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
echo "bench rc=$?" >> gpurun_out/bench_b.err
grep -E "passed|failed" gpurun_out/pytest.txt | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest.txt | head -20; tail -2 gpurun_out/ncu_solve.log; wc -l gpurun_out/trace.txt
