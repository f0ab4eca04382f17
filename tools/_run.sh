timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest.log
