#!/bin/bash
mkdir -p gpurun_out
for v in expw expws; do
DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_$v.so timeout 200 python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline --frames 400 > gpurun_out/h_$v.json 2> gpurun_out/h_$v.err
echo "$v rc=$? $(cut -c56-80 gpurun_out/h_$v.json)"; tail -1 gpurun_out/h_$v.err
done
