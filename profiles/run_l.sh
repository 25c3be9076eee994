#!/bin/bash
# records with the three-object kernel (1 GPU) + A/B of the tensor-core K build of the correspondence branch
mkdir -p gpurun_out
python profiles/run_train.py 8 73 0.1 2>/dev/null | grep "N=8" > gpurun_out/l_cls_fp32.txt
DCD_B200_KTC=1 python profiles/run_train.py 8 73 0.1 2>/dev/null | grep "N=8" > gpurun_out/l_cls_ktc.txt
echo "fp32 K:"; cat gpurun_out/l_cls_fp32.txt; echo "tcgen05 K:"; cat gpurun_out/l_cls_ktc.txt
timeout 900 python bench.py --config sweep1m --steps 2 --warmup 3 --graph-collective > gpurun_out/bench_sweep1m_1gpu.json 2> gpurun_out/bench_sweep1m.err
echo "sweep rc=$? $(cut -c1-160 gpurun_out/bench_sweep1m_1gpu.json)"
timeout 600 python bench.py --config kitti_val_full --steps 2 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_kitti_val_full_1gpu.json 2> gpurun_out/bench_full.err
echo "full rc=$? $(cut -c1-160 gpurun_out/bench_kitti_val_full_1gpu.json)"
