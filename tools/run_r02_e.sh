B="python bench.py --steps 3 --warmup 3 --only"
for v in default b8 sw b8sw b2; do
  if [ $v = default ]; then unset PCREG_LIB; else export PCREG_LIB=/root/repo/ab/libpcreg_$v.so; fi
  $B > gpurun_out/b_$v.json 2>/dev/null
  echo "== $v"; python tools/bench_brief.py gpurun_out/b_$v.json
done
for v in default b8; do
  if [ $v = default ]; then unset PCREG_LIB; else export PCREG_LIB=/root/repo/ab/libpcreg_$v.so; fi
  echo "== c5 $v"; python tools/c5_check.py 256 2>&1 | grep "voxel map profiling 2"
done
