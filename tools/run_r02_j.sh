#!/bin/bash
# round 2, call j: the final tree -- whole GPU suite, smoke, default bench line, reference arm
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/j_tests.log
cat gpurun_out/j_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py ) > gpurun_out/j_bench_default.json 2> gpurun_out/j_bench_default.err
tail -4 gpurun_out/j_bench_default.err
( time python bench.py --impl reference ) > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err
tail -4 gpurun_out/j_bench_ref.err
python tools/bench_brief.py gpurun_out/j_bench_default.json
