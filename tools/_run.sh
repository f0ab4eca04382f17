set -x
python -m pytest tests/test_gpu_staged.py tests/test_gpu_parity_pixels.py -x -q -m gpu -s 2>&1 | grep -E "config 1|scores,|similarity .auto|candidate lists|passed|failed|Error" > gpurun_out/r2_pytest10.log
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -x -q -m gpu -k "similarity or config2" 2>&1 | tail -3 >> gpurun_out/r2_pytest10.log
cat gpurun_out/r2_pytest10.log
python bench.py --steps 2 --warmup 1 --headline-only --one-arm --no-cpu-baseline > gpurun_out/r2_plain_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c.csv python bench.py --steps 2 --warmup 1 --headline-only --one-arm --no-cpu-baseline > gpurun_out/r2_ncu_c.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -c 1500 gpurun_out/r2_bench_c.json; tail -3 gpurun_out/r2_bench_c.err
