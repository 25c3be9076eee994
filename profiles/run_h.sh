#!/bin/bash
# GPU pass of the two-objects-per-group fused forward: parity of the fused kernel, a short bench, the in-kernel timeline
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gmw.py -m gpu -x -q -k "fused_inference_forward or full_size" > gpurun_out/h_pytest1.txt 2>&1
echo "pytest1 rc=$?" | tee -a gpurun_out/h_pytest1.txt
tail -3 gpurun_out/h_pytest1.txt
if grep -q "pytest1 rc=0" gpurun_out/h_pytest1.txt; then
  timeout 300 python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
  echo "bench rc=$?"; cat gpurun_out/h_bench.json | cut -c1-330
  DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_trace.so TRACE_OBJECTS=8 timeout 120 python profiles/trace_fused.py > gpurun_out/h_trace.txt 2>&1
  echo "trace rc=$?"; wc -l gpurun_out/h_trace.txt
fi
