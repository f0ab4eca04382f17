python tools/prof_cnn.py 296 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 5 -c 1 -o gpurun_out/prof_cnn3 python tools/prof_cnn.py 296 > gpurun_out/ncu_prof_cnn.log 2>&1
tail -2 gpurun_out/ncu_prof_cnn.log
