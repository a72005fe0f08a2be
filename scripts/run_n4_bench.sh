cd /root/repo
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/r2_bench_n4.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
        ls=d.get('large_scene')
        if ls: print('large', ls['value'], ls['ms_per_step'], ls['ms_fwd_bwd'])
PY
