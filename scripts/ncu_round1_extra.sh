#!/bin/bash
# ncu --set full of (a) the memory-bound render kernels, (b) the training backward kernels.  One GPU; outputs under gpurun_out/.
set -x
K1='regex:knn_kernel|ray_select_kernel|ray_compact_kernel|project_views_kernel|nbr_weights_kernel|image_gather_fwd_v2_kernel|blend_fwd_kernel|composite_fwd_kernel|linear_fwd_smalln_kernel'
timeout 600 ncu --set full --clock-control none -k "$K1" -c 24 -f -o gpurun_out/prof_membound python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-train > gpurun_out/ncu_mb.log 2>&1
K2='regex:linear_tc_kernel|wgrad_tc_kernel|bwd|adam_kernel'
timeout 600 ncu --set full --clock-control none -k "$K2" -c 70 -f -o gpurun_out/prof_trainbwd python scripts/train_step_bench.py --steps 1 --warmup 0 > gpurun_out/ncu_tb.log 2>&1
ls -la gpurun_out/*.ncu-rep
