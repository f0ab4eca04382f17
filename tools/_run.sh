set -x
python -m pytest tests/test_gpu_parity_pixels.py tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -x -q -m gpu -s 2>&1 | tail -60 > gpurun_out/r2_pytest2.log
tail -30 gpurun_out/r2_pytest2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 6000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
