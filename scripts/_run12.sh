cd /root/repo
timeout 200 python scripts/dp_timeline.py > gpurun_out/dp_timeline_n1.txt 2>&1; echo "rc=$?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/dp_timeline_n1.txt | head -16
timeout 200 python scripts/cpu_issue_time.py 2>&1 | tail -4
