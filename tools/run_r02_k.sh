#!/bin/bash
# round 2, call k20: fused kernel, trimmed mode: no gather for the correspondences outside the selection (sums pass) A/B
set -x
timeout 900 python -m pytest tests/test_gpu_voxel_map.py tests/test_gpu_icp.py -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --only"
V=pcreg_b200/variants/libpcreg_noskip.so
$B > gpurun_out/k20_skip.json 2>/dev/null
PCREG_LIB=$V $B > gpurun_out/k20_noskip.json 2>/dev/null
$B > gpurun_out/k20_skip2.json 2>/dev/null
PCREG_LIB=$V $B > gpurun_out/k20_noskip2.json 2>/dev/null
python tools/bench_brief.py gpurun_out/k20_skip.json gpurun_out/k20_noskip.json gpurun_out/k20_skip2.json gpurun_out/k20_noskip2.json
