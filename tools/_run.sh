timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -s -k "cta_pair" > gpurun_out/pytest_pair.log 2>&1; echo "exit $?"; grep "pair kernel\|passed\|failed\|Error" gpurun_out/pytest_pair.log | tail -14
timeout 600 python tools/sweep_encoder.py > gpurun_out/sweep_encoder.log 2>&1; tail -6 gpurun_out/sweep_encoder.log
