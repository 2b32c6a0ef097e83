python -m pytest tests/test_gpu_configs.py tests/test_gpu_kabsch_ransac.py -x -q > gpurun_out/t_cfg.log 2>&1; tail -15 gpurun_out/t_cfg.log
( time python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_full.err
python tools/bench_brief.py gpurun_out/bench_full.json
