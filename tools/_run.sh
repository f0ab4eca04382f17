timeout 600 python -m pytest tests -x -q -m gpu -k "cnn or conv or hamming or config3" 2>&1 | tail -3
timeout 200 python tools/bench_cnnvtl.py 2>&1 | tail -1 | cut -c1-200
DLC_DEBUG_SET="6=1" timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm -s 5 -c 5 --csv --log-file /tmp/l.csv python tools/prof_cnn.py 1063 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10]
h=rows[0]; i_k=h.index('Kernel Name'); i_m=h.index('Metric Name'); i_v=h.index('Metric Value'); i_id=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[i_id], r[i_k][5:50]),{})[r[i_m][:12]]=r[i_v]
for k,v in d.items(): print(k, v)
PY
