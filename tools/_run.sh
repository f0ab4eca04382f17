timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bench_v8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
