set -x
KREG='regex:gemm_pair_kernel|gram_refine_fix_kernel|prep_rows_kernel|colsum_partial_kernel|gram_probe_kernel|patch_gather|topk_rows|weights_centre|rep_from_hash'
for w in normal xavier; do
python bench.py --steps 1 --warmup 1 --headline-only --one-arm --no-cpu-baseline --weights $w > gpurun_out/r2_plainfull_$w.log 2>&1 && \
ncu --set full --clock-control none -k "$KREG" -c 15 -o /tmp/r2_prof_$w python bench.py --steps 1 --warmup 1 --headline-only --one-arm --no-cpu-baseline --weights $w > gpurun_out/r2_ncufull_$w.log 2>&1
ncu -i /tmp/r2_prof_$w.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_$w.csv 2>/dev/null
ncu -i /tmp/r2_prof_$w.ncu-rep --page details > gpurun_out/r2_ncu_details_$w.txt 2>/dev/null
done
python bench.py --config 4 > gpurun_out/r2_bench_config4.json 2>gpurun_out/r2_bench_config4.err
du -sh gpurun_out
