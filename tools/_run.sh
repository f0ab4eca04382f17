timeout 900 python -m pytest tests -x -q -m gpu -k "patch or smoke or config2 or sda or stream" > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['ms_per_launch'], d['stages_ms'])"
python -c "import __graft_entry__ as g; g.smoke()"
