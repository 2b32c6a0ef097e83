# round-2 evidence: launch list + --set full captures of the kernels bench.py's roofline objects name
set -x
A="python bench.py --steps 1 --warmup 3 --only"
B="python bench.py --workload c5 --steps 1 --warmup 1 --only"
C="python bench.py --steps 1 --warmup 3 --only --workload c2"
$A > gpurun_out/plainA.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3.csv $A > gpurun_out/ncuA0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_icp_fused -s 3 -c 1 -o gpurun_out/r02_full_c3_fused -f $A > gpurun_out/ncuA1.log 2>&1
$B > gpurun_out/plainB.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_nn_vox -s 30 -c 2 -o gpurun_out/r02_full_c5_vox -f $B > gpurun_out/ncuB1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_icp_update -s 30 -c 1 -o gpurun_out/r02_full_c5_update -f $B > gpurun_out/ncuB2.log 2>&1
$C > gpurun_out/plainC.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_nn_brute -s 40 -c 1 -o gpurun_out/r02_full_c2_brute -f $C > gpurun_out/ncuC1.log 2>&1
D="python tools/local_points_run.py"
$D > gpurun_out/plainD.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_local_grid -s 1 -c 1 -o gpurun_out/r02_full_local_grid -f $D > gpurun_out/ncuD1.log 2>&1
ls -la gpurun_out | tail -20
