#!/bin/bash
# round 2, call k13: k_local_grid with per-lane forward stepping instead of a bisection per candidate: tests, timing, ncu capture
set -x
timeout 900 python -m pytest tests/test_gpu_local_points.py tests/test_gpu_descriptors.py tests/test_gpu_pipeline.py -x -q -s 2>&1 | tail -4
D="python tools/local_points_run.py"
$D > gpurun_out/plainD.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_local_grid -s 1 -c 1 -o gpurun_out/r02_full_local_grid -f $D > gpurun_out/ncuD1.log 2>&1
cat gpurun_out/plainD.log
