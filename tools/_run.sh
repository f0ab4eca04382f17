set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_pytest_final.log; cat gpurun_out/r2_pytest_final.log
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 600 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2>&1; tail -c 400 gpurun_out/r2_bench_ref_final.json
python bench.py --steps 2 --warmup 1 --headline-only --one-arm --no-cpu-baseline > gpurun_out/r2_plain_d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_d.csv python bench.py --steps 2 --warmup 1 --headline-only --one-arm --no-cpu-baseline > gpurun_out/r2_ncu_d.log 2>&1
