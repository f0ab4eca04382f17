timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['ms_per_launch'], d['stages_ms'])"
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from deeploopcloser_b200 import ops
S=torch.randn(1063,1063,device='cuda')
for _ in range(3): ops.topk_rows(S,10,exclude_band=0)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): ops.topk_rows(S,10,exclude_band=0)
e1.record(); torch.cuda.synchronize(); print('topk_rows 1063x1063 k=10: %.1f us'%(e0.elapsed_time(e1)/50*1e3))
PY
