#!/bin/bash
# ncu --set full of the fused forward (one launch, 2048 objects), summarised on the box
mkdir -p gpurun_out/r02
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:mlp_fused_kernel -c 1 -o gpurun_out/r02/mlp_fused_kernel python profiles/run_gmw_infer.py 2048 1 > gpurun_out/r02/mlp_fused_kernel.log 2>&1
python profiles/summarise.py kernel gpurun_out/r02/mlp_fused_kernel.ncu-rep gpurun_out/r02/r02_mlp_fused_kernel.md mlp_fused_kernel >> gpurun_out/r02/mlp_fused_kernel.log 2>&1
python profiles/summarise.py traffic gpurun_out/r02/mlp_fused_kernel.ncu-rep mlp_fused_kernel mlp_fused_kernel > gpurun_out/r02/traffic.log 2>&1
ncu -i gpurun_out/r02/mlp_fused_kernel.ncu-rep --page source --csv > gpurun_out/r02/fused_source.csv 2>/dev/null
rm -f gpurun_out/r02/*.ncu-rep
tail -3 gpurun_out/r02/mlp_fused_kernel.log; ls -la gpurun_out/r02 | head
