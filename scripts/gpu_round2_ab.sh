#!/bin/bash
# First GPU call of round 2: A/B of the switches prepared (but not measured) at the end of round 1.  One B200, ~4 min.
#   HNR_KNN_V3=1            third-generation kNN kernel (query.cu)            -> must stay bit-exact (query suites), query stage ms
#   HNR_TC_BWD_MIN_N=65     SIMT backward for layers <= 64 wide (ops.py)      -> train step ms, gradient parity suites
#   HNR_MAX_VALID_CHUNK=N   valid samples per aggregator pass (default 262144) -> render ms per frame
set -u
run() { echo "== $*"; env "$@" 2>&1 | tail -3; }
HNR_KNN_V3=1 timeout 150 python -m pytest tests/test_gpu_query.py tests/test_gpu_query_vs_reference.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/ab_knn_v3_tests.log 2>&1; echo "knn v3 parity rc=$?"; tail -2 gpurun_out/ab_knn_v3_tests.log
HNR_TC_BWD_MIN_N=65 timeout 150 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_aggregator.py tests/test_gpu_basic_ops.py -m gpu -x -q > gpurun_out/ab_bwd_simt_tests.log 2>&1; echo "skinny SIMT backward parity rc=$?"; tail -2 gpurun_out/ab_bwd_simt_tests.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/ab_render_default.json 2>/dev/null
HNR_KNN_V3=1 timeout 100 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/ab_render_knn_v3.json 2>/dev/null
HNR_MAX_VALID_CHUNK=524288 timeout 100 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/ab_render_chunk512k.json 2>/dev/null
HNR_MAX_VALID_CHUNK=131072 timeout 100 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/ab_render_chunk128k.json 2>/dev/null
timeout 100 python scripts/train_step_bench.py --steps 5 --warmup 3 --json gpurun_out/ab_train_default.json > /dev/null 2>&1
HNR_TC_BWD_MIN_N=65 timeout 100 python scripts/train_step_bench.py --steps 5 --warmup 3 --json gpurun_out/ab_train_bwd_simt.json > /dev/null 2>&1
timeout 60 python scripts/measure_tf32_peak.py --json gpurun_out/tf32_peak.json
python - <<'P'
import glob, json
for f in sorted(glob.glob("gpurun_out/ab_render_*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["ms_per_step"], 2), "ms/frame", d["roofline"]["stage_ms"])
    except Exception as e:
        print(f, "unreadable", e)
for f in sorted(glob.glob("gpurun_out/ab_train_*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["ms_fwd_bwd"], 2), "ms fwd+bwd", {k: v for k, v in d["stage_ms"].items() if "linear" in k})
    except Exception as e:
        print(f, "unreadable", e)
P
