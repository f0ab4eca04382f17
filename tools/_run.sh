for i in 1 2; do
for L in "" "/root/repo/deeploopcloser_b200/libdlc_ab.so"; do
echo "== lib=$L"
DLC_LIB_PATH=$L timeout 300 python bench.py --headline-only --one-arm --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms'].items()})"
DLC_LIB_PATH=$L timeout 200 python tools/bench_cnnvtl.py 2>&1 | tail -1 | cut -c1-130
done; done
