timeout 300 python tools/ab_layer_bk.py > gpurun_out/ab_layer_bk.log 2>&1; tail -9 gpurun_out/ab_layer_bk.log
