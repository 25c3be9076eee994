#!/bin/bash
# Round-2 profile captures (one GPU).  Launch list of the bench command, then ncu --set full of the top kernels.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r02.log 2>&1
timeout 600 $NCU -k regex:mlp_fused_kernel -c 1 -o gpurun_out/prof_mlp_fused_r02 python profiles/run_gmw_infer.py 2048 1 > gpurun_out/ncu_fused.log 2>&1
timeout 600 $NCU -k regex:'edge_mean_block|edge_select_radix|edge_solve_bwd' -c 4 -o gpurun_out/prof_solve_r02 python profiles/run_solve.py > gpurun_out/ncu_solve.log 2>&1
timeout 900 $NCU -k regex:'tb_|transport_' -s 20 -c 40 -o gpurun_out/prof_transport_r02 python profiles/run_train.py 8 73 0.1 > gpurun_out/ncu_transport.log 2>&1
timeout 900 $NCU -k regex:'mlp_tc_kernel|mlp_bwd_tc_kernel' -s 60 -c 12 -o gpurun_out/prof_train_r02 python profiles/run_train.py 8 73 0 > gpurun_out/ncu_train.log 2>&1
timeout 900 $NCU -k regex:'mlp_tc_kernel' -s 4 -c 3 -o gpurun_out/prof_mlp_tc_n256_r02 python profiles/run_train.py 16 256 0 > gpurun_out/ncu_n256.log 2>&1
python profiles/run_train.py 8 73 0.1 > gpurun_out/train_cls_times.txt 2>&1
python profiles/run_train.py 8 73 0 >> gpurun_out/train_cls_times.txt 2>&1
ls -la gpurun_out/*.ncu-rep; cat gpurun_out/train_cls_times.txt; tail -2 gpurun_out/launches_r02.log
