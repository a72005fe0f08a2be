#!/bin/bash
# launch list of the FINAL training step (after the first-layer split of the blend-weight net, the one-pass head backward and the vector
# reductions) + ncu --set full of the kernels added last; one GPU, reports land in gpurun_out/
set -u
T=scripts/train_step_bench.py
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_train.csv python $T --steps 1 --warmup 2 > gpurun_out/ncu_r2g_tl.log 2>&1; echo "train list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:chain_f16_kernel|linear_head_bwd|img_sum_views|image_gather_bwd_v2|alpha_ksum_bwd_img" --launch-skip 36 -c 8 -o gpurun_out/r2g_train_kernels -f python $T --steps 1 --warmup 2 > gpurun_out/ncu_r2g_a.log 2>&1; echo "kernels rc=$?"
