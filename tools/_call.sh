mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c5.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/ab17_parity.log 2>&1
for c in c3 c5; do timeout 300 python bench.py --config $c --no-cpu > gpurun_out/ab17_$c.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/ab17_$c.json').read().strip().splitlines()[-1]); print('$c fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])"; done
cat gpurun_out/ab17_parity.log
