set -x
timeout 300 python tools/diag_cnnvtl.py > gpurun_out/diag_cnn.log 2>&1
tail -40 gpurun_out/diag_cnn.log
timeout 300 python tools/bench_cnnvtl.py 128 > gpurun_out/cnn_fused_128.log 2>&1; tail -3 gpurun_out/cnn_fused_128.log
timeout 300 python tools/bench_cnnvtl.py 1063 > gpurun_out/cnn_fused_1063.log 2>&1; tail -3 gpurun_out/cnn_fused_1063.log
