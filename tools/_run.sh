timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
timeout 300 python tools/bench_cnnvtl.py 1063 > gpurun_out/cnn_fused_1063.log 2>&1; tail -1 gpurun_out/cnn_fused_1063.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['ms_per_launch'], d['stages_ms'])"
python - <<'PY'
import sys; sys.path.insert(0,'.')
from deeploopcloser_b200 import _lib
import subprocess
PY
