#!/bin/bash
# final validation of round 1 on one B200: GPU suite (batched kNN walk = default), A/B of the kNN walk on the same box, bench
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/t27.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/t27.log | cut -c1-300
if [ $rc -ne 0 ]; then
  grep -n "Error\|assert" gpurun_out/t27.log | head -20 | cut -c1-300
  HNR_KNN_PER_VOXEL=1 timeout 150 python -m pytest tests -m gpu -q > gpurun_out/t27b.log 2>&1; echo "pytest(per-voxel kNN, no -x) rc=$?"; tail -6 gpurun_out/t27b.log | cut -c1-300
fi
timeout 150 python bench.py > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "bench rc=$?"; tail -2 gpurun_out/bench21.err
HNR_KNN_PER_VOXEL=1 timeout 100 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/bench21b.json 2> gpurun_out/bench21b.err; echo "bench(per-voxel kNN) rc=$?"
python - <<'P'
import json
for f in ("gpurun_out/bench21.json", "gpurun_out/bench21b.json"):
    try:
        d = json.load(open(f))
        print(f, round(d["ms_per_step"], 2), "ms", d["roofline"]["stage_ms"], d.get("train", {}).get("ms_fwd_bwd"))
    except Exception as e:
        print(f, "unreadable", e)
P
