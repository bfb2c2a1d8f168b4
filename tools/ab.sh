#!/bin/bash
# A/B of build variants on the GPU box: tools/ab.sh "" _t32x16 ...   (suffixes of draw_b200/libdraw_b200<suffix>.so)
# Prints parity (subset) and the bench's frames/s + per-kernel ms for c3 c2 c4 (c5 with AB_C5=1).
for v in "$@"; do
  export DRAW_B200_LIB=/root/repo/draw_b200/libdraw_b200$v.so
  echo "== variant '$v'"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3 or c1_textured or odd or c4_dungeon_flythrough_4k or degenerate or ties" 2>&1 | tail -1
  for c in c3 c2 c4 ${AB_C5:+c5}; do python bench.py --config $c --steps ${AB_STEPS:-100} --warmup 10 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$c', 'fps', round(d['value'],1), 'us', round(1e3*d['ms_per_step'],1), 'flushed_us', round(1e3*d['config']['ms_per_step_l2_flushed'],1), 'e2e', round(d['e2e']['value'],1), {k:round(1e3*v,1) for k,v in d['kernel_ms'].items()})"; done
done
