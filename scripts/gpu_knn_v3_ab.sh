#!/bin/bash
# A/B of the third-generation kNN kernel (csrc/query.cu: knn_kernel_v3, off by default).  One B200, ~1.5 min.
#   1. bit-exactness: the query parity suites (numpy oracle + the reference's own pycuda kernels) with HNR_KNN_V3=1
#   2. speed: render bench stage split with and without it on the same box
HNR_KNN_V3=1 timeout 120 python -m pytest tests/test_gpu_query.py tests/test_gpu_query_vs_reference.py tests/test_gpu_e2e.py -m gpu -x -q > gpurun_out/knn_v3_tests.log 2>&1
echo "knn v3 parity rc=$?"; tail -3 gpurun_out/knn_v3_tests.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/knn_ab_default.json 2>/dev/null
HNR_KNN_V3=1 timeout 100 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/knn_ab_v3.json 2>/dev/null
python - <<'P'
import json
for f in ("gpurun_out/knn_ab_default.json", "gpurun_out/knn_ab_v3.json"):
    d = json.load(open(f))
    print(f, round(d["ms_per_step"], 2), "ms/frame, query", d["roofline"]["stage_ms"]["query"], "ms")
P
