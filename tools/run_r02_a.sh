python -m pytest tests/test_gpu_voxel_map.py tests/test_gpu_icp.py -x -q > gpurun_out/t1.log 2>&1; tail -12 gpurun_out/t1.log
B="python bench.py --steps 3 --warmup 3 --no-c2 --no-match --no-cpu"
$B > gpurun_out/b_fused.log 2>&1
PCREG_FUSED=0 $B > gpurun_out/b_unfused.log 2>&1
for v in t384 t512b1 t256; do PCREG_LIB=/root/repo/ab/libpcreg_$v.so $B > gpurun_out/b_$v.log 2>&1; done
python tools/bench_brief.py gpurun_out/b_fused.log gpurun_out/b_unfused.log gpurun_out/b_t384.log gpurun_out/b_t512b1.log gpurun_out/b_t256.log
tail -3 gpurun_out/b_fused.log | cut -c1-600
