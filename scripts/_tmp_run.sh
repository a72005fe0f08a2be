cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --no-large --no-extras --no-cpu-baseline > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.err; echo "bench rc=$?"
timeout 100 python scripts/bench_chain.py 2>&1 | grep chain
