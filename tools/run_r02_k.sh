#!/bin/bash
# round 2, call k14: descriptor stage on a model with a grid: test + timing
set -x
timeout 900 python -m pytest tests/test_gpu_descriptors.py -x -q 2>&1 | tail -4
python tools/desc_run.py 20000 2>&1 | tail -4
