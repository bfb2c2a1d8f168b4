mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "painter or transparent or c1_" 2>&1 | tail -15 ) > gpurun_out/ab7_parity.log 2>&1
( DRAW_B200_LIB=/root/repo/draw_b200/libdraw_b200_t256.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3 or c1_textured or odd or stripes or c4_dungeon_flythrough_4k or ties" 2>&1 | tail -3 ) >> gpurun_out/ab7_parity.log 2>&1
L=/root/repo/draw_b200/libdraw_b200
tools/gpu_ab.sh "A=0" "DRAW_B200_LIB=${L}_t256.so" "DRAW_B200_LIB=${L}_t256.so DRAW_B200_TILE_CTAS=592" > gpurun_out/ab7.log 2>&1
cat gpurun_out/ab7_parity.log gpurun_out/ab7.log
