#!/bin/bash
# round 2, call k21: getLocalPoints test incl. the >32768-hit fallback of the grid path
set -x
timeout 900 python -m pytest tests/test_gpu_local_points.py -x -q -s 2>&1 | tail -4
