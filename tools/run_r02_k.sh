#!/bin/bash
# round 2, call k18: k_icp_update on C5: 512 threads x 1 block (128 regs) vs 512 x 2 (64 regs, spills) vs 256 x 3 (80 regs)
set -x
B="python bench.py --steps 2 --warmup 1 --only --workload c5"
$B > gpurun_out/k18_base.json 2>/dev/null
PCREG_LIB=pcreg_b200/variants/libpcreg_mb2.so $B > gpurun_out/k18_mb2.json 2>/dev/null
PCREG_LIB=pcreg_b200/variants/libpcreg_nt256.so $B > gpurun_out/k18_nt256.json 2>/dev/null
python tools/bench_brief.py gpurun_out/k18_*.json
