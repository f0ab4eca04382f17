timeout 900 python -m pytest tests -x -q -m gpu -k "match or stream or topk" > gpurun_out/pytest.log 2>&1; tail -3 gpurun_out/pytest.log
timeout 300 python tools/bench_matcher.py --batches 1,32,128,256,512,1024,4096 > gpurun_out/matcher.log 2>&1; grep '^{' gpurun_out/matcher.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['B'], round(d['ms'],3), 'ms', round(d['tflops'],1), 'TF/s', round(d['tflops_frac_of_measured_sustained'],3), round(d['db_gbs'],1), 'GB/s', round(d['hbm_frac_of_measured'],3), d['planted_top1_ok'])
"
