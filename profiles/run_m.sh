#!/bin/bash
# A/B: last statistics merge on the converter threads (-DDCD_FUSED_CONV_FINAL=1) vs on the statistics warps
mkdir -p gpurun_out
DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_cf.so timeout 300 python -m pytest tests/test_gpu_gmw.py -m gpu -x -q -k "fused or full_size" > gpurun_out/m_pytest_cf.txt 2>&1
echo "pytest cf rc=$?"; tail -2 gpurun_out/m_pytest_cf.txt
DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_cf.so timeout 200 python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --frames 400 > gpurun_out/m_cf.json 2> gpurun_out/m_cf.err
echo "cf rc=$? $(cut -c56-80 gpurun_out/m_cf.json)"
timeout 200 python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --frames 400 > gpurun_out/m_base.json 2> gpurun_out/m_base.err
echo "base rc=$? $(cut -c56-80 gpurun_out/m_base.json)"
