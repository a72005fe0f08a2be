cd /root/repo
timeout 60 python scripts/bench_chain.py > gpurun_out/r5_chain.log 2>&1; echo "bench_chain rc=$?"; grep chain gpurun_out/r5_chain.log
timeout 240 python -m pytest tests/test_gpu_mlp_tc.py tests/test_gpu_aggregator.py tests/test_gpu_fused_bwd.py tests/test_gpu_e2e.py -x -q --timeout 60 > gpurun_out/r5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r5_pytest.log
timeout 200 python scripts/train_step_bench.py > gpurun_out/r2_train_i.json 2> gpurun_out/r2_train_i.err; echo "train rc=$?"; tail -2 gpurun_out/r2_train_i.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_train_i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_fwd_bwd'], d['host_issue_ms_per_step'], d['e2e']['ms_per_step']); print(d['stage_ms'])"
