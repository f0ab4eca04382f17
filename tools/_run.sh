for A in normal xavier; do
python bench.py --steps 1 --warmup 1 --headline-only --one-arm --no-cpu-baseline --weights $A > gpurun_out/r2_plain_$A.log 2>&1 || exit 1
timeout 500 ncu --set full --clock-control none --import-source on -c 40 -o /tmp/full_$A python bench.py --steps 1 --warmup 1 --headline-only --one-arm --no-cpu-baseline --weights $A > gpurun_out/r2_ncu_full_$A.log 2>&1
ncu -i /tmp/full_$A.ncu-rep --page raw --csv > gpurun_out/r2_full_raw_$A.csv 2>/dev/null
ls -la /tmp/full_$A.ncu-rep gpurun_out/r2_full_raw_$A.csv
done
