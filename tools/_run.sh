timeout 300 python tools/diag_cnnvtl.py > gpurun_out/diag_cnn.log 2>&1
grep "descriptors\|FAILED\|Error" gpurun_out/diag_cnn.log | tail; grep "conv1 " gpurun_out/diag_cnn.log
timeout 900 python -m pytest tests -x -q -m gpu -k "cnnvtl" > gpurun_out/pytest_cnn.log 2>&1; tail -5 gpurun_out/pytest_cnn.log
timeout 300 python tools/bench_cnnvtl.py 128 > gpurun_out/cnn_fused_128.log 2>&1; tail -1 gpurun_out/cnn_fused_128.log
timeout 300 python tools/bench_cnnvtl.py 1063 > gpurun_out/cnn_fused_1063.log 2>&1; tail -1 gpurun_out/cnn_fused_1063.log
