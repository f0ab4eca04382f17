set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "encoder or sda" 2>&1 | tail -5
timeout 200 python tools/bench_overlap_1gpu.py --weights normal > gpurun_out/r2_overlap_normal.json 2> gpurun_out/r2_overlap_normal.err; cat gpurun_out/r2_overlap_normal.json; tail -3 gpurun_out/r2_overlap_normal.err
timeout 200 python tools/bench_overlap_1gpu.py --weights xavier > gpurun_out/r2_overlap_xavier.json 2> gpurun_out/r2_overlap_xavier.err; cat gpurun_out/r2_overlap_xavier.json; tail -3 gpurun_out/r2_overlap_xavier.err
timeout 200 python bench.py --config 5 --no-cpu-baseline > gpurun_out/r2_config5_wave.json 2> gpurun_out/r2_config5_wave.err; tail -c 700 gpurun_out/r2_config5_wave.json; tail -3 gpurun_out/r2_config5_wave.err
