#!/bin/bash
# round 2, call h: grid-based getLocalPoints -- tests + the bench leg
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_local_points.py tests/test_gpu_descriptors.py -x -q -s 2>&1 | tail -15 > gpurun_out/h_tests.log
timeout 600 python bench.py --only --no-cpu 2>gpurun_out/h_bench.err | tail -1 > gpurun_out/h_bench_only.json
timeout 900 python - > gpurun_out/h_local.json 2>gpurun_out/h_local.err <<'PY'
import json, torch, bench
import pcreg_b200 as P
P.init(0)
print(json.dumps(bench.local_points_workload(P, torch), indent=1))
PY
cat gpurun_out/h_tests.log; cat gpurun_out/h_local.json; tail -3 gpurun_out/h_local.err
