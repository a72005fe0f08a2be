cd /root/repo
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench_final3.json 2> gpurun_out/r2_bench_final3.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
S=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref3.json 2> gpurun_out/r2_bench_ref3.err; echo "ref rc=$? wall=$(( $(date +%s) - S ))s"
