mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/ab14_parity.log 2>&1
( env DRAW_B200_CLEAR_IN_TILE=3 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -x -q 2>&1 | tail -3 ) >> gpurun_out/ab14_parity.log 2>&1
AB_CFGS="c3 c2 c4" tools/gpu_ab.sh "A=0" "DRAW_B200_CLEAR_IN_TILE=3" "DRAW_B200_CLEAR_IN_TILE=3 DRAW_B200_TILE_CTAS=444" "DRAW_B200_CLEAR_CTAS=296" "DRAW_B200_CLEAR_IN_TILE=1" "DRAW_B200_CLEAR_IN_TILE=2" > gpurun_out/ab14.log 2>&1
AB_CFGS="c5" AB_STEPS=60 tools/gpu_ab.sh "A=0" "DRAW_B200_CLEAR_IN_TILE=3" >> gpurun_out/ab14.log 2>&1
cat gpurun_out/ab14_parity.log gpurun_out/ab14.log
