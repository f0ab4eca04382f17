timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest.log
timeout 300 python tools/bench_streaming.py --db-rows 1000000 --batch 256 --steps 5 > gpurun_out/stream1.log 2>&1; grep '^{' gpurun_out/stream1.log | tail -1
timeout 300 python tools/bench_streaming.py --db-rows 1000000 --batch 256 --steps 5 --detect > gpurun_out/stream1d.log 2>&1; grep '^{' gpurun_out/stream1d.log | tail -1
