timeout 900 python -m pytest tests -x -q -m gpu -k "surf or detects" > gpurun_out/pytest_surf.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_surf.log
timeout 600 python tools/bench_surf.py > gpurun_out/bench_surf.log 2>&1; cat gpurun_out/bench_surf.log | tail -5 | cut -c1-200
