#!/bin/bash
# round 2, call i3: list scan tail handling (clamped index + select vs predicated load of a far pair) A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_voxel_map.py tests/test_gpu_icp.py -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --only"
V=pcreg_b200/variants/libpcreg_noclamp.so
$B > gpurun_out/i_clamp.json 2>/dev/null
PCREG_LIB=$V $B > gpurun_out/i_noclamp.json 2>/dev/null
$B > gpurun_out/i_clamp2.json 2>/dev/null
PCREG_LIB=$V $B > gpurun_out/i_noclamp2.json 2>/dev/null
python tools/bench_brief.py gpurun_out/i_clamp.json gpurun_out/i_noclamp.json gpurun_out/i_clamp2.json gpurun_out/i_noclamp2.json
python tools/c4_check.py 16384 2>&1 | tail -1
PCREG_LIB=$V python tools/c4_check.py 16384 2>&1 | tail -1
