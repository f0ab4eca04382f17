for lib in libdlc_old.so libdlc_noelect.so libdlc.so libdlc_old.so; do
  echo "== $lib"
  DLC_LIB_PATH=$PWD/deeploopcloser_b200/$lib timeout 300 python tools/bench_matcher.py --batches 32,1024 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['B'], round(d['ms'],3), round(d['tflops'],1), round(d['db_gbs'],1))
    elif 'rror' in l: print(l.strip()[:200])
"
done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv
