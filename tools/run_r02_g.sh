python -m pytest tests/test_gpu_voxel_map.py tests/test_gpu_icp.py tests/test_gpu_configs.py -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --only"
$B > gpurun_out/b_skip.json 2>/dev/null; PCREG_FUSED_SKIP=0 $B > gpurun_out/b_noskip.json 2>/dev/null
python tools/bench_brief.py gpurun_out/b_skip.json gpurun_out/b_noskip.json
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/b_skip.json') if l.startswith('{')][0]
print(d['roofline']['phase_share'], d['roofline'].get('entries_per_query'))
"
python tools/fused_phases.py 2>&1 | tail -1
python tools/c4_check.py 16384 2>&1 | tail -2
PCREG_FUSED_SKIP=0 python tools/c4_check.py 16384 2>&1 | tail -1
