#!/bin/bash
# One gpurun call: smoke, parity tests, bench lines, launch list, ncu full captures.  Outputs under gpurun_out/.
mkdir -p gpurun_out
T=${TAG:-r02}
( timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/${T}_smoke.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
for c in c2 c4 c5; do timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_c3_reference.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 90 --csv --log-file gpurun_out/${T}_c3_launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tile|k_raster|k_front' -s 300 -c 3 -f -o gpurun_out/${T}_prof_c3 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${T}_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_tile' -s 40 -c 1 -f -o gpurun_out/${T}_prof_c4_tile python bench.py --config c4 --steps 2 --warmup 3 --quick > gpurun_out/${T}_ncu_c4.log 2>&1
cat gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_bench_c3.json
