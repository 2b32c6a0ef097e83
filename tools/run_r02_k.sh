#!/bin/bash
# round 2, call k16: AlignPoints kernels: histogram selection + unrolled streaming loops vs HEAD (median of three launches)
set -x
F="--steps 2 --warmup 3 --no-c2 --no-cpu --no-match --no-c4 --no-c5 --no-local"
for r in 1 2; do
python bench.py $F > gpurun_out/k16_new$r.json 2>/dev/null
PCREG_LIB=pcreg_b200/variants/libpcreg_head.so python bench.py $F > gpurun_out/k16_head$r.json 2>/dev/null
done
python - <<'PY'
import json
for f in ('new1','head1','new2','head2'):
    d=[json.loads(l) for l in open('gpurun_out/k16_%s.json'%f) if l.startswith('{')][0]
    print(f, {k:(round(v['kernel_ms'],3), round(v['roofline']['frac'],3)) for k,v in d['align_batch']['variants'].items()})
PY
