TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 300 $TR tools/ab_sharded_seq.py > gpurun_out/r2_ab_sharded_8gpu.log 2>&1; grep '^{' gpurun_out/r2_ab_sharded_8gpu.log > gpurun_out/r2_ab_sharded_8gpu.jsonl; cut -c48-230 gpurun_out/r2_ab_sharded_8gpu.jsonl; grep -E "Error|Traceback" gpurun_out/r2_ab_sharded_8gpu.log | head -3
BEST=$(python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2_ab_sharded_8gpu.jsonl')]
rows=[r for r in rows if r['weights']=='normal' and r['wave_pad']==1 and r['pipeline_stages']==2]
print(min(rows,key=lambda r:r['ms_per_sequence'])['sm_reserve'] if rows else 0)
PY
)
echo "best reserve $BEST"
DLC_SM_RESERVE=$BEST timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_8gpu_b.log 2>&1; echo "exit $?"; tail -1 gpurun_out/r2_bench_8gpu_b.log > gpurun_out/r2_bench_8gpu_b.json; cut -c1-400 gpurun_out/r2_bench_8gpu_b.json
