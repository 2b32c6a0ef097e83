#!/bin/bash
# round 2, call k19: k_icp_update<512> on C5 with 6 / 8 correspondences per trip instead of 4
set -x
B="python bench.py --steps 2 --warmup 1 --only --workload c5"
$B > gpurun_out/k19_ub4.json 2>/dev/null
PCREG_LIB=pcreg_b200/variants/libpcreg_ub6.so $B > gpurun_out/k19_ub6.json 2>/dev/null
PCREG_LIB=pcreg_b200/variants/libpcreg_ub8.so $B > gpurun_out/k19_ub8.json 2>/dev/null
python tools/bench_brief.py gpurun_out/k19_*.json
