python -m pytest tests/test_gpu_voxel_map.py tests/test_gpu_nn.py -x -q 2>&1 | tail -2
B="python bench.py --steps 3 --warmup 3 --only"
for v in default g1 g2 g8; do
  if [ $v = default ]; then unset PCREG_LIB; else export PCREG_LIB=/root/repo/ab/libpcreg_$v.so; fi
  $B > gpurun_out/b_$v.json 2>/dev/null
  PCREG_FUSED=0 $B > gpurun_out/bu_$v.json 2>/dev/null
  echo "== $v"; python tools/bench_brief.py gpurun_out/b_$v.json gpurun_out/bu_$v.json
  python tools/c5_check.py 128 2>&1 | grep "voxel map profiling 2"
done
