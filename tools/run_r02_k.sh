#!/bin/bash
# round 2, call k9: k_icp_update with the correspondence indices loaded one trip ahead (+ batched key / select passes) A/B vs HEAD
set -x
timeout 900 python -m pytest tests/test_gpu_voxel_map.py tests/test_gpu_icp.py -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --only"
V=pcreg_b200/variants/libpcreg_head.so
$B --workload c5 > gpurun_out/k9_c5_new.json 2>/dev/null
PCREG_LIB=$V $B --workload c5 > gpurun_out/k9_c5_head.json 2>/dev/null
PCREG_FUSED=0 $B > gpurun_out/k9_c3pp_new.json 2>/dev/null
PCREG_FUSED=0 PCREG_LIB=$V $B > gpurun_out/k9_c3pp_head.json 2>/dev/null
$B --workload c2g > gpurun_out/k9_c2g_new.json 2>/dev/null
PCREG_LIB=$V $B --workload c2g > gpurun_out/k9_c2g_head.json 2>/dev/null
python tools/bench_brief.py gpurun_out/k9_*.json
