#!/bin/bash
# One gpurun call: smoke, the GPU tests and the bench lines of C2-C5 with the driver's flags + the reference arm (no ncu).
mkdir -p gpurun_out
T=${TAG:-r02}
( timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/${T}_smoke.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
for c in c2 c4 c5; do timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_c3_reference.json 2>/dev/null
cat gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_pytest.log
python - <<PY
import json
for c in ("c2","c3","c4","c5"):
    d=json.loads(open("gpurun_out/${T}_bench_%s.json" % c).read().strip().splitlines()[-1]); e=d["e2e"]
    print(c, round(d["value"],1), round(d["us_per_frame"],2), "e2e", round(e["value"],1), round(e["serial_value"],1), "MB/frame", round(e["d2h_bytes_per_step"]/d["config"]["frames_per_step"]/1e6,2), "frac", round(d["roofline"]["frac"],3))
PY
