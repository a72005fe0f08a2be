cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r6_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r6_pytest.log
