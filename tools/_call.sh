mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/ab11_parity.log 2>&1
L=/root/repo/draw_b200/libdraw_b200
tools/gpu_ab.sh "A=0" "DRAW_BENCH_SHARED_STREAM=1" > gpurun_out/ab11.log 2>&1
AB_CFGS="c3" tools/gpu_ab.sh "DRAW_B200_CLEAR_IN_TILE=1" "DRAW_B200_CLEAR_IN_TILE=3" "DRAW_B200_BIN_RPW=0" "DRAW_B200_BIN_RPW=0 DRAW_B200_CLIP_CTAS=74" "DRAW_B200_SETS=6" "DRAW_B200_SETS=8" "DRAW_B200_KPRIO=1" "DRAW_B200_KPRIO=1 DRAW_B200_SETS=6" "DRAW_B200_COST_SHADE=590 DRAW_B200_SPLIT_DIV=1024 DRAW_B200_SPLIT_MIN_COST=128" "DRAW_B200_LIB=${L}_t256.so" "DRAW_B200_LIB=${L}_t256.so DRAW_B200_SETS=6" "DRAW_B200_SETS=6 DRAW_B200_BIN_RPW=0 DRAW_B200_CLEAR_IN_TILE=3" >> gpurun_out/ab11.log 2>&1
cat gpurun_out/ab11_parity.log gpurun_out/ab11.log
