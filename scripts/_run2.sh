cd /root/repo
python -m pytest tests/test_gpu_mlp_tc.py tests/test_gpu_aggregator.py tests/test_gpu_fused_bwd.py tests/test_gpu_e2e.py -x -q 2>&1 | tail -6
python scripts/bench_chain.py 2>&1 | tail -8
python scripts/trace_chain.py cf 250 > gpurun_out/trace_cf2.txt 2>&1
python scripts/trace_chain.py am 250 > gpurun_out/trace_am2.txt 2>&1
head -2 gpurun_out/trace_cf2.txt gpurun_out/trace_am2.txt
