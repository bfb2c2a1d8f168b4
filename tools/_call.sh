mkdir -p gpurun_out
L=/root/repo/draw_b200/libdraw_b200
( timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -x -q -k "c3 or c2 or c1_ or odd or stripes or ties or export or c4_dungeon_flythrough_4k" 2>&1 | tail -3 ) > gpurun_out/ab18_parity.log 2>&1
( env DRAW_B200_LIB=${L}_rpipe.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c3 or c2 or c1_ or odd or stripes or ties" 2>&1 | tail -3 ) >> gpurun_out/ab18_parity.log 2>&1
AB_CFGS="c3 c2 c4" tools/gpu_ab.sh "A=0" "DRAW_B200_LIB=${L}_nopipe.so" "DRAW_B200_LIB=${L}_rpipe.so" "A=1" "DRAW_B200_LIB=${L}_nopipe.so" > gpurun_out/ab18.log 2>&1
cat gpurun_out/ab18_parity.log gpurun_out/ab18.log
