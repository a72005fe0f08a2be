cd /root/repo
timeout 400 python -m pytest tests/test_gpu_basic_ops.py tests/test_gpu_aggregator.py tests/test_gpu_fused_bwd.py tests/test_gpu_e2e.py -x -q --timeout 100 2>&1 | tail -3
timeout 300 python scripts/train_step_bench.py --steps 10 > gpurun_out/r2_train_k.json 2> gpurun_out/r2_train_k.err; echo "train rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_train_k.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_fwd_bwd'], d['host_issue_ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step']); print(d['stage_ms'])"
