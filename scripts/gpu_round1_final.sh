#!/bin/bash
# final validation of round 1 on one B200: GPU suite, smoke, bench, then one ncu --set full capture of the batched kNN kernel
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/t28.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t28.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 170 python bench.py > gpurun_out/bench22.json 2> gpurun_out/bench22.err; echo "bench rc=$?"; tail -2 gpurun_out/bench22.err
python - <<'P'
import json
d = json.load(open("gpurun_out/bench22.json"))
print(round(d["ms_per_step"], 2), "ms", d["roofline"]["stage_ms"], d["train"]["ms_fwd_bwd"], d["cpu_baseline"])
P
timeout 75 ncu --set full --clock-control none --import-source on -k regex:knn_kernel -c 1 -f -o gpurun_out/prof_knn_batched python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-train > gpurun_out/ncu_knn.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_knn_batched.ncu-rep
