#!/bin/bash
# round 2, call h: grid-based getLocalPoints -- the bench leg + the whole GPU suite on this tree
set -x
mkdir -p gpurun_out
timeout 900 python - > gpurun_out/h_local.json 2>gpurun_out/h_local.err <<'PY'
import json, torch, bench
import pcreg_b200 as P
P.init(0)
print(json.dumps(bench.local_points_workload(P, torch), indent=1))
PY
cat gpurun_out/h_local.json; tail -3 gpurun_out/h_local.err
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/h_tests_all.log
cat gpurun_out/h_tests_all.log
