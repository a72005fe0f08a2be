cd /root/repo
timeout 300 python scripts/train_step_bench.py --steps 10 > gpurun_out/r2_train_j.json 2> gpurun_out/r2_train_j.err; echo "train rc=$?"; tail -2 gpurun_out/r2_train_j.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_train_j.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_fwd_bwd'], d['host_issue_ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step']); print(d['stage_ms'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_train3.csv python scripts/train_step_bench.py --steps 3 --warmup 3 > /dev/null 2>&1; echo "ncu rc=$?"
