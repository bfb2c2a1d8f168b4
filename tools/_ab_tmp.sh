for d in 3 4 6 8; do
  DRAW_BENCH_E2E_DEPTH=$d timeout 400 python bench.py --config c3 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('depth $d c3', 'fps', round(d['value'],1), 'e2e', round(e['value'],1), 'serial', round(e['serial_value'],1), 'achieved', round(e['d2h_achieved_gbs'],2))"
done
DRAW_BENCH_E2E_DEPTH=6 timeout 400 python bench.py --config c2 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('depth 6 c2', 'e2e', round(e['value'],1), 'serial', round(e['serial_value'],1), 'achieved', round(e['d2h_achieved_gbs'],2))"
DRAW_BENCH_E2E_DEPTH=6 timeout 400 python bench.py --config c4 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('depth 6 c4', 'e2e', round(e['value'],1), 'serial', round(e['serial_value'],1), 'achieved', round(e['d2h_achieved_gbs'],2))"
