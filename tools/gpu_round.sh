#!/bin/bash
# One gpurun call: parity tests, bench lines, launch list, ncu full captures.  Outputs under gpurun_out/.
mkdir -p gpurun_out
T=${TAG:-r01e}
( timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2 ) > gpurun_out/${T}_smoke.log
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${T}_pytest.log
timeout 300 python bench.py > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
for c in c2 c4 c5; do timeout 300 python bench.py --config $c --no-cpu > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_c3_reference.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/${T}_c3_launches.csv python bench.py --steps 30 --warmup 10 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_tile|k_raster|k_setup' -s 54 -c 3 -f -o gpurun_out/${T}_prof_c3 python bench.py --steps 20 --warmup 10 --no-cpu > gpurun_out/${T}_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_tile' -s 40 -c 1 -f -o gpurun_out/${T}_prof_c4_tile python bench.py --config c4 --steps 10 --warmup 3 --quick > gpurun_out/${T}_ncu_c4.log 2>&1
cat gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_bench_c3.json
