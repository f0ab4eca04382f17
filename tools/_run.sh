for i in 1 2; do
for L in "" "/root/repo/deeploopcloser_b200/libdlc_old.so"; do
echo "== lib=$L"
DLC_LIB_PATH=$L timeout 200 python tools/bench_matcher.py --batches 1024 --reps 8 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k in ('B','ms','tflops','batch')})"
done; done
