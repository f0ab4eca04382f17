timeout 900 python -m pytest tests -x -q -m gpu -k "surf" 2>&1 | tail -6
timeout 300 python tools/bench_surf.py 2>&1 | tail -3 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none -k regex:surf_octave -s 8 -c 4 --csv --log-file /tmp/l.csv python tools/bench_surf.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('/tmp/l.csv')) if len(r)>10]
h=rows[0]; i_k=h.index('Kernel Name'); i_m=h.index('Metric Name'); i_v=h.index('Metric Value'); i_id=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[i_id], r[i_k][:40]),{})[r[i_m][:34]]=r[i_v]
for k,v in d.items(): print(k, v)
PY
