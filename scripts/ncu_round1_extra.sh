#!/bin/bash
# ncu counters of the memory-bound render kernels (few metrics = few replay passes; --set full on 24 launches of a process holding
# several GB did not finish in 10 minutes) and the launch list of one training step.  One GPU; small csv outputs under gpurun_out/.
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
K1='regex:knn_kernel|ray_select_kernel|ray_compact_kernel|project_views_kernel|nbr_weights_kernel|image_gather_fwd_v2_kernel|blend_fwd_kernel|composite_fwd_kernel|linear_fwd_smalln_kernel'
timeout 150 ncu --metrics $M --clock-control none -k "$K1" -c 14 --csv --log-file gpurun_out/membound_r1.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-train > gpurun_out/ncu_mb.log 2>&1
echo "membound rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_r1.csv python scripts/train_step_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_tl.log 2>&1
echo "train list rc=$?"
K2='regex:alpha_ksum_bwd|linear_bwd_data_narrow|linear_bwd_weight_smalln|image_gather_kernel|nbr_features|conf_up_bwd|composite_bwd|blend_bwd'
timeout 100 ncu --metrics $M --clock-control none -k "$K2" -c 10 --csv --log-file gpurun_out/membound_train_r1.csv python scripts/train_step_bench.py --steps 1 --warmup 0 > gpurun_out/ncu_tm.log 2>&1
echo "train membound rc=$?"
ls -la gpurun_out/*.csv
