#!/bin/bash
# Round-2 profile captures (one GPU).  Launch list of the bench command, then ncu --set full of the top kernels; the reports
# are summarised ON the box (profiles/summarise.py needs only the ncu CLI) and deleted: gpurun_out/ must stay below 64 MiB.
mkdir -p gpurun_out/r02
NCU="ncu --set full --clock-control none --import-source on -f"
S="python profiles/summarise.py"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > gpurun_out/r02/launches_r02.log 2>&1
$S launches gpurun_out/r02/launches_r02.csv gpurun_out/r02/r02_launches.md
cap() {  # cap <name> <kernel-regex> <extra ncu args...> -- <command...>
  name=$1; regex=$2; shift; shift
  args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
  timeout 900 $NCU -k regex:"$regex" "${args[@]}" -o gpurun_out/r02/$name "$@" > gpurun_out/r02/$name.log 2>&1
  $S kernel gpurun_out/r02/$name.ncu-rep gpurun_out/r02/r02_$name.md "$regex" >> gpurun_out/r02/$name.log 2>&1
}
cap mlp_fused_kernel 'mlp_fused_kernel' -c 1 -- python profiles/run_gmw_infer.py 2048 1
$S traffic gpurun_out/r02/mlp_fused_kernel.ncu-rep mlp_fused_kernel mlp_fused_kernel > gpurun_out/r02/traffic.log 2>&1
cap edge_solve 'edge_mean_block|edge_select_radix|edge_solve_bwd' -c 4 -- python profiles/run_solve.py
cap transport 'tb_feat|tb_row|tb_col|tb_cg|transport_k|transport_row|transport_col' -s 30 -c 14 -- python profiles/run_train.py 8 73 0.1
cap train 'mlp_tc_kernel|mlp_bwd_tc_kernel' -s 60 -c 8 -- python profiles/run_train.py 8 73 0
cap mlp_tc_n256 'mlp_tc_kernel' -s 4 -c 2 -- python profiles/run_train.py 16 256 0
$S traffic gpurun_out/r02/mlp_tc_n256.ncu-rep mlp_tc_kernel_n256 mlp_tc_kernel >> gpurun_out/r02/traffic.log 2>&1
cp profiles/traffic.json gpurun_out/r02/traffic.json 2>/dev/null
rm -f gpurun_out/r02/*.ncu-rep
python profiles/run_train.py 8 73 0.1 > gpurun_out/r02/train_cls_times.txt 2>&1
python profiles/run_train.py 8 73 0 >> gpurun_out/r02/train_cls_times.txt 2>&1
du -sh gpurun_out; ls gpurun_out/r02; cat gpurun_out/r02/traffic.log
