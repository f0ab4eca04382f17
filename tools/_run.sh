set -x
python -m pytest tests/test_gpu_training.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest3.log
tail -8 gpurun_out/r2_pytest3.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 3000 gpurun_out/r2_bench_b.json; tail -5 gpurun_out/r2_bench_b.err
