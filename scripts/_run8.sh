cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2c.json 2> gpurun_out/r2_bench_n2c.err; echo "rc=$?"
tail -c 600 gpurun_out/r2_bench_n2c.err
python - <<'PY'
import json
for line in open('gpurun_out/r2_bench_n2c.json'):
    if line.startswith('{'):
        d=json.loads(line)
        print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
        print(d['train'].get('ranks'))
        ls=d.get('large_scene')
        if ls: print('large', ls['value'], ls['ms_per_step'], ls['ms_fwd_bwd'], ls.get('ranks'))
PY
