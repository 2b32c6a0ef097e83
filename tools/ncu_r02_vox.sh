set -x
export PCREG_LANES=1
CMD="python bench.py --steps 1 --warmup 3 --no-c2 --no-match --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_nn_vox -s 0 -c 2 -o gpurun_out/r02_vox_first -f $CMD > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_nn_vox -s 25 -c 2 -o gpurun_out/r02_vox_steady -f $CMD > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_icp_update -s 25 -c 2 -o gpurun_out/r02_update -f $CMD > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log
ls -la gpurun_out/
