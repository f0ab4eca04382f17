timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest.log; grep "raw-pixel encoder" gpurun_out/pytest.log
for f in "" "--split-pixel-input"; do
timeout 900 python bench.py --no-cpu-baseline $f > gpurun_out/bench_ab$f.log 2>&1; tail -1 gpurun_out/bench_ab$f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['ms_per_launch'], d['stages_ms'])"
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_ab2.log 2>&1; tail -1 gpurun_out/bench_ab2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['ms_per_launch'], d['stages_ms'])"
