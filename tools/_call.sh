mkdir -p gpurun_out
L=/root/repo/draw_b200/libdraw_b200
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/ab12_parity.log 2>&1
tools/gpu_ab.sh "A=0" "DRAW_B200_SETS=5" "DRAW_B200_SETS=6" "DRAW_B200_LIB=${L}_t256.so DRAW_B200_SETS=6" "DRAW_B200_LIB=${L}_t256.so DRAW_B200_SETS=6 DRAW_B200_TILE_CTAS=444" > gpurun_out/ab12.log 2>&1
env DRAW_B200_SETS=6 python tools/trace_frames.py c3 24 own > gpurun_out/trace_c3_own6.log 2>&1
cat gpurun_out/ab12_parity.log gpurun_out/ab12.log; tail -14 gpurun_out/trace_c3_own6.log
