#!/bin/bash
# round 2, call k22: model creation breakdown on C5 / C4 / C3 (per-level voxel build log)
set -x
for c in c5 c4 c3; do python tools/vox_build_times.py $c 2>&1 | grep -E "pcreg vox|create" | cut -c1-220; done
