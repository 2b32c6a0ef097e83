python -m pytest tests/test_gpu_voxel_map.py -x -q -k fused > gpurun_out/t1.log 2>&1; tail -3 gpurun_out/t1.log
python tools/fused_phases.py
for v in t384 t512b1; do echo $v; PCREG_LIB=/root/repo/ab/libpcreg_$v.so python tools/fused_phases.py; done
