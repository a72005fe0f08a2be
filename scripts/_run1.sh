set -x
cd /root/repo
python -m pytest tests/test_gpu_fused_bwd.py -x -q 2>&1 | tail -5
python scripts/train_step_bench.py > gpurun_out/r2_train_h.json 2> gpurun_out/r2_train_h.err; tail -3 gpurun_out/r2_train_h.err
HNR_SIDE_STREAM=0 python scripts/train_step_bench.py > gpurun_out/r2_train_h_noside.json 2>/dev/null
python scripts/trace_chain.py cf 300 > gpurun_out/trace_cf.txt 2>&1
python scripts/trace_chain.py am 300 > gpurun_out/trace_am.txt 2>&1
python scripts/trace_chain.py cm 300 > gpurun_out/trace_cm.txt 2>&1
