cd /root/repo
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 scripts/check_peer_allreduce.py > gpurun_out/peer_check_n2.txt 2>&1; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/peer_check_n2.txt | tail -25
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 scripts/dp_timeline.py > gpurun_out/dp_timeline_n2.txt 2>&1; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/dp_timeline_n2.txt | head -16
