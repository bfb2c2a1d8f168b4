timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -x -q 2>&1 | tail -3
AB_CFGS="c4 c3 c5" AB_STEPS=100 bash tools/gpu_ab.sh "DRAW_B200_SORT_LARGE=1" "DRAW_B200_SORT_LARGE=1"
