timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -x -q 2>&1 | tail -3
for r in 1 2; do
for c in c3 c4 c5 c2; do
    python bench.py --config $c --steps 100 --warmup 20 --quick 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  $c fps', round(d['value'],1), 'us', round(d['us_per_frame'],2), 'lone', round(d['lone_frame_us_median'],1))"
done; done
