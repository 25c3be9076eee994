#!/bin/bash
mkdir -p gpurun_out
for c in trace trace6 oldtrace; do
DCD_B200_LIB=$PWD/dcd_b200/libdcd_b200_$c.so TRACE_OBJECTS=8 timeout 120 python profiles/trace_fused.py > gpurun_out/h_$c.txt 2>&1
echo "$c rc=$?"; wc -l gpurun_out/h_$c.txt
done
