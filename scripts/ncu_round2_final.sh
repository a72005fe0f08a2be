#!/bin/bash
# Round-2 final evidence (one GPU; reports land in gpurun_out/): launch lists of the final training step and rendered frame, and
# ncu --set full captures of the kernels changed after scripts/ncu_round2.sh ran: chain_f16 (eight worker warps), wgrad_img (3-stage
# ring), image_gather_bwd_v2 (vector reductions), plus the two big tensor kernels of the step for reference.
set -u
T=scripts/train_step_bench.py
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_train.csv python $T --steps 1 --warmup 2 > gpurun_out/ncu_r2f_tl.log 2>&1; echo "train list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:chain_f16_kernel|image_gather_bwd_v2|wgrad_img_kernel|nbr_mlp_f16_kernel|nbr_bwd_f16_kernel|chain_bwd_f16" --launch-skip 26 -c 13 -o gpurun_out/r2f_train_kernels -f python $T --steps 1 --warmup 2 > gpurun_out/ncu_r2f_a.log 2>&1; echo "train kernels rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_render.csv python bench.py --steps 3 --warmup 3 --no-large --no-extras --no-cpu-baseline > gpurun_out/ncu_r2f_rl.log 2>&1; echo "render list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:chain_f16_kernel" --launch-skip 3 -c 3 -o gpurun_out/r2f_chain_render -f python scripts/bench_chain.py > gpurun_out/ncu_r2f_b.log 2>&1; echo "chain render rc=$?"
ls -la gpurun_out/r2f_*
