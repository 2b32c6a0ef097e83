#!/bin/bash
# usage: tools/sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   -> one short C3 bench per setting
for cfg in "$@"; do
  out=$(env $cfg python bench.py --steps 2 --warmup 3 --no-c2 --no-cpu --no-match 2>&1 | tail -1)
  echo "$cfg :: $(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step %.1f  Gq/s %.3f  nn_ms %.1f upd_ms %.1f' % (d['ms_per_step'], d['value']/1e9, d['kernel_time_share']['nn_ms'], d['kernel_time_share']['update_ms']))")"
done
