timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q -s > gpurun_out/pytest_full.log 2>&1; tail -25 gpurun_out/pytest_full.log
