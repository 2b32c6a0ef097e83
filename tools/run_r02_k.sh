#!/bin/bash
# round 2, call k17: descriptor kernel with the one-pass histogram selection: tests + timing
set -x
timeout 900 python -m pytest tests/test_gpu_descriptors.py tests/test_gpu_pipeline.py tests/test_gpu_golden_rows.py -x -q 2>&1 | tail -3
python tools/desc_run.py 20000 2>&1 | tail -3
