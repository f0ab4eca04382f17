set -x
python -m pytest tests/test_gpu_staged.py -x -q -m gpu -s 2>&1 | tail -25 > gpurun_out/r2_pytest4.log
tail -25 gpurun_out/r2_pytest4.log
for w in normal xavier; do
python bench.py --steps 2 --warmup 1 --headline-only --one-arm --no-cpu-baseline --weights $w > gpurun_out/r2_plain_$w.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_$w.csv python bench.py --steps 2 --warmup 1 --headline-only --one-arm --no-cpu-baseline --weights $w > gpurun_out/r2_ncu_$w.log 2>&1
done
ls -la gpurun_out/r2_launches_*.csv
