#!/bin/bash
mkdir -p gpurun_out
for v in expw exps expws; do
DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_$v.so timeout 200 python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --frames 800 > gpurun_out/h_$v.json 2> gpurun_out/h_$v.err
echo "$v rc=$? $(cut -c1-120 gpurun_out/h_$v.json)"; tail -2 gpurun_out/h_$v.err
done
timeout 200 python bench.py --steps 1 --warmup 3 --quick --no-cpu-baseline --frames 800 > gpurun_out/h_base.json 2> gpurun_out/h_base.err
echo "base rc=$? $(cut -c1-120 gpurun_out/h_base.json)"
