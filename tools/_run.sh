timeout 300 python tools/ab_probe_side.py > gpurun_out/ab_probe_side.log 2>&1; tail -3 gpurun_out/ab_probe_side.log
timeout 900 python -m pytest tests -x -q -m gpu -k "simil or sdav or fullsize or pipeline or stream" > gpurun_out/pytest_sim.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_sim.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_nocpu.log 2>&1; tail -1 gpurun_out/bench_nocpu.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], d['stages_ms'])"
