python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -o gpurun_out/prof_gram_v2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gram.log 2>&1
tail -2 gpurun_out/ncu_gram.log
python tools/bench_matcher.py --rows 500000 --batches 1024 --reps 3 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/prof_match_v2 python tools/bench_matcher.py --rows 500000 --batches 1024 --reps 3 > gpurun_out/ncu_match2.log 2>&1
tail -1 gpurun_out/ncu_match2.log
