timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_full4.log 2>&1; tail -4 gpurun_out/r2_pytest_full4.log
