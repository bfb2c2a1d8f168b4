#!/bin/bash
# A/B on the GPU box: env-knob variants, quick bench (timed region only).  Usage: tools/gpu_ab.sh "VAR=val VAR2=val" ...
q() { python bench.py --config $1 --steps ${AB_STEPS:-300} --warmup 30 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  $1 fps', round(d['value'],1), 'us', round(d['us_per_frame'],2), 'lone us median/max', round(d['lone_frame_us_median'],1), round(d['lone_frame_us_max'],1))"; }
for v in "$@"; do
  echo "== $v"
  for c in ${AB_CFGS:-c3 c2 c4}; do env $v bash -c "$(declare -f q); q $c"; done
done
