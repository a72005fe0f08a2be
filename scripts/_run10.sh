cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 scripts/dp_timeline.py > gpurun_out/dp_timeline_n8.txt 2>&1; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/dp_timeline_n8.txt | head -60
