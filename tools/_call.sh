mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_setup|k_raster|k_bin|k_vertex' -s 36 -c 5 -f -o gpurun_out/r01_prof_c5 python bench.py --config c5 --steps 6 --warmup 3 --quick > gpurun_out/r01_ncu_c5.log 2>&1
tail -3 gpurun_out/r01_ncu_c5.log | cut -c1-300
AB_CFGS="c5" AB_STEPS=60 tools/gpu_ab.sh "A=0" "DRAW_B200_SETS=4" >> gpurun_out/ab16.log 2>&1
AB_CFGS="c3 c2 c4" tools/gpu_ab.sh "A=0" >> gpurun_out/ab16.log 2>&1
cat gpurun_out/ab16.log
