#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gmw.py -m gpu -x -q -k "fused or full_size" > gpurun_out/h_pytest1.txt 2>&1
echo "pytest1 rc=$?"; tail -2 gpurun_out/h_pytest1.txt
for v in vol; do
DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_$v.so timeout 200 python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --frames 400 > gpurun_out/h_$v.json 2> gpurun_out/h_$v.err
echo "$v rc=$? $(cut -c56-80 gpurun_out/h_$v.json)"
done
timeout 200 python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --frames 400 > gpurun_out/h_base.json 2> gpurun_out/h_base.err
echo "base(gpu scope) rc=$? $(cut -c56-80 gpurun_out/h_base.json)"
