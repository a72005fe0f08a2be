#!/bin/bash
# Round-2 evidence: launch lists (gpu__time_duration per launch) of one training step and one rendered frame, and ncu --set full
# captures of the kernels that carry the round: the fused backward (nbr_bwd_f16, wgrad_img, chain_bwd_f16), the training forward
# (nbr_mlp_f16<2>, chain_f16) and the per-sample chains of the render path.  One GPU; reports land in gpurun_out/.
set -u
T=scripts/train_step_bench.py
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train.csv python $T --steps 1 --warmup 2 > gpurun_out/ncu_r2_tl.log 2>&1; echo "train list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nbr_bwd_f16_kernel --launch-skip 2 -c 1 -o gpurun_out/r2_nbr_bwd_f16 -f python $T --steps 1 --warmup 2 > gpurun_out/ncu_r2_a.log 2>&1; echo "nbr_bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_img_kernel --launch-skip 8 -c 4 -o gpurun_out/r2_wgrad_img -f python $T --steps 1 --warmup 2 > gpurun_out/ncu_r2_b.log 2>&1; echo "wgrad_img rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain_bwd_f16_kernel --launch-skip 6 -c 3 -o gpurun_out/r2_chain_bwd -f python $T --steps 1 --warmup 2 > gpurun_out/ncu_r2_c.log 2>&1; echo "chain_bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:nbr_mlp_f16_kernel|chain_f16_kernel" --launch-skip 8 -c 4 -o gpurun_out/r2_train_fwd -f python $T --steps 1 --warmup 2 > gpurun_out/ncu_r2_d.log 2>&1; echo "train fwd rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_render.csv python bench.py --steps 3 --warmup 3 --no-large --no-extras --no-cpu-baseline > gpurun_out/ncu_r2_rl.log 2>&1; echo "render list rc=$?"
ls -la gpurun_out/r2_*.ncu-rep gpurun_out/r2_launches_*.csv
