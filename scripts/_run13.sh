cd /root/repo
timeout 200 python scripts/dp_timeline.py > gpurun_out/dp_timeline_n1b.txt 2>&1; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/dp_timeline_n1b.txt | head -8
timeout 200 python -m pytest tests/test_gpu_fused_bwd.py -x -q --timeout 100 2>&1 | tail -2
