timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c5.py tests/test_gpu_paths.py -x -q 2>&1 | tail -5
for c in c5 c3 c4 c2; do
  python bench.py --config $c --steps 60 --warmup 10 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  $c fps', round(d['value'],1), 'us', round(d['us_per_frame'],2), 'lone', round(d['lone_frame_us_median'],1))"
done
python tools/list_stats.py c5 2>&1 | grep "k_front phases" | head -1
