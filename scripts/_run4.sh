cd /root/repo
ncu --set full --clock-control none --import-source on -k regex:chain_f16 -s 3 -c 1 -f -o gpurun_out/r2_chain_cf_v2 python scripts/bench_chain.py > gpurun_out/ncu_chain_v2.log 2>&1
tail -3 gpurun_out/ncu_chain_v2.log
