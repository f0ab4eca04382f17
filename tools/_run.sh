timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "detector" > gpurun_out/pytest_fs.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_fs.log
