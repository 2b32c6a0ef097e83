#!/bin/bash
# round 2, call k10: list scan with whole trips peeled off (no predicates / far pairs in them) A/B
set -x
timeout 900 python -m pytest tests/test_gpu_voxel_map.py tests/test_gpu_icp.py -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --only"
V=pcreg_b200/variants/libpcreg_nopeel.so
$B > gpurun_out/k10_peel.json 2>/dev/null
PCREG_LIB=$V $B > gpurun_out/k10_nopeel.json 2>/dev/null
$B > gpurun_out/k10_peel2.json 2>/dev/null
PCREG_LIB=$V $B > gpurun_out/k10_nopeel2.json 2>/dev/null
python tools/bench_brief.py gpurun_out/k10_peel.json gpurun_out/k10_nopeel.json gpurun_out/k10_peel2.json gpurun_out/k10_nopeel2.json
python tools/c4_check.py 16384 2>&1 | tail -1
PCREG_LIB=$V python tools/c4_check.py 16384 2>&1 | tail -1
