timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/pytest_train.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_train.log
