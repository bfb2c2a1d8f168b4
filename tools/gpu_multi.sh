#!/bin/bash
# One gpurun --gpus N call: the bench line at N GPUs (frame-parallel value, e2e, sort-first entries) and the multi-GPU tests.
# Usage (on the GPU box): N=8 TAG=r02 bash tools/gpu_multi.sh
N=${N:-2}; T=${TAG:-r02}
mkdir -p gpurun_out
timeout 1300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_c3_n$N.json 2> gpurun_out/${T}_bench_n$N.err
tail -c 600 gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_c3_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","us_per_frame","n_gpus")}, "e2e", d["e2e"]["value"], d["e2e"]["serial_value"], d["e2e"]["d2h_ceiling_gbs"])
for name, sf in (d.get("sort_first") or {}).items():
    print(name, "single", sf.get("single_gpu_ms_per_frame"), "lone", sf.get("single_gpu_lone_frame_ms"))
    for k, v in sf.items():
        if isinstance(v, dict) and "ms_per_frame" in v:
            print("   ", k, round(v["ms_per_frame"], 4), "x", round(v["speedup_vs_single_gpu"], 2), "bit_exact", v["bit_exact"])
PY
( timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 )
