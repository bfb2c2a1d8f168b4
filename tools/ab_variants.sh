for v in "" _t32 _t64a; do
  export DRAW_B200_LIB=/root/repo/draw_b200/libdraw_b200$v.so
  echo "== variant '$v'"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3 or c1_textured or odd or stripes or c4_dungeon_flythrough_4k" 2>&1 | tail -1
  for c in c3 c2 c4; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$c', 'fps', round(d['value'],1), 'ms', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['kernel_ms'].items()})"; done
done
