#!/bin/bash
# like sweep.sh but 8 timed steps per setting and only the device-path step time (less sensitive to host hiccups)
for cfg in "$@"; do
  out=$(env $cfg python bench.py --steps 8 --warmup 3 --no-c2 --no-cpu --no-match 2>&1 | tail -1)
  echo "$cfg :: $(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step %.1f  e2e %.1f  Gq/s %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['value']/1e9))")"
done
