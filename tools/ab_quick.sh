#!/bin/bash
# tools/ab_quick.sh <configs> -- <lib suffixes>: bench only, no parity
cfgs=(); while [ "$1" != "--" ]; do cfgs+=("$1"); shift; done; shift
for v in "$@"; do
  export DRAW_B200_LIB=/root/repo/draw_b200/libdraw_b200$v.so
  echo "== variant '$v'"
  for c in "${cfgs[@]}"; do python bench.py --config $c --steps ${AB_STEPS:-200} --warmup 20 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$c', 'fps', round(d['value'],1), 'us', round(1e3*d['ms_per_step'],1), 'flushed_us', round(1e3*d['config']['ms_per_step_l2_flushed'],1), 'e2e', round(d['e2e']['value'],1), {k:round(1e3*v,1) for k,v in d['kernel_ms'].items()})"; done
done
