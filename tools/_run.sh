timeout 900 python -m pytest tests/test_gpu_training.py -x -q -s > gpurun_out/pytest_train.log 2>&1; tail -30 gpurun_out/pytest_train.log
