#!/bin/bash
# round 2, call k7: prefetch of the next trip's voxel headers in the fused kernel A/B
set -x
timeout 900 python -m pytest tests/test_gpu_voxel_map.py tests/test_gpu_icp.py -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --only"
V=pcreg_b200/variants/libpcreg_nopf.so
$B > gpurun_out/k7_pf.json 2>/dev/null
PCREG_LIB=$V $B > gpurun_out/k7_nopf.json 2>/dev/null
$B > gpurun_out/k7_pf2.json 2>/dev/null
PCREG_LIB=$V $B > gpurun_out/k7_nopf2.json 2>/dev/null
python tools/bench_brief.py gpurun_out/k7_pf.json gpurun_out/k7_nopf.json gpurun_out/k7_pf2.json gpurun_out/k7_nopf2.json
python tools/c4_check.py 16384 2>&1 | tail -1
PCREG_LIB=$V python tools/c4_check.py 16384 2>&1 | tail -1
python tools/fused_phases.py 2>&1 | tail -1
