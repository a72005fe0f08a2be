cd /root/repo
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 scripts/dp_timeline.py > gpurun_out/dp_timeline_n8b.txt 2>&1; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/dp_timeline_n8b.txt | head -14
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 scripts/check_peer_allreduce.py > gpurun_out/peer_check_n8.txt 2>&1; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/peer_check_n8.txt | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 10 --warmup 3 --no-large > gpurun_out/r2_bench_n8d.json 2> gpurun_out/r2_bench_n8d.err; echo "rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/r2_bench_n8d.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
        print([(r['ms_fwd_bwd_no_collective'], r['host_issue_ms_per_step']) for r in d['train'].get('ranks')])
PY
