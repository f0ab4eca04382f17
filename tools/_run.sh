timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-400
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:GramPolicy -s 1 -c 1 -o gpurun_out/prof_gram_v2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gram.log 2>&1
tail -2 gpurun_out/ncu_gram.log
