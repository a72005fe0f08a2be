cd /root/repo
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2_bench_final1.json 2> gpurun_out/r2_bench_final1.err; echo "rc=$? wall=$(( $(date +%s) - S ))s"
tail -c 400 gpurun_out/r2_bench_final1.err
S=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref1.json 2> gpurun_out/r2_bench_ref1.err; echo "ref rc=$? wall=$(( $(date +%s) - S ))s"
cut -c1-600 gpurun_out/r2_bench_ref1.json
