cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
timeout 300 python scripts/train_step_bench.py --steps 10 > gpurun_out/r2_train_n.json 2> gpurun_out/r2_train_n.err; echo "train rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_train_n.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_fwd_bwd'], d['host_issue_ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'])"
