python -m pytest tests -m gpu -x -q -k "knn" > gpurun_out/t_knn.log 2>&1; tail -1 gpurun_out/t_knn.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn_ --csv --log-file gpurun_out/l.csv python tools/one_forward.py > gpurun_out/ncu_a.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/l.csv') if not l.startswith('=='))]
h=rows[0];vi=h.index("Metric Value");ki=h.index("Kernel Name")
for n in ('knn_pack','knn_tc_kernel','knn_finish'):
    print(n,[float(r[vi])/1e3 for r in rows[1:] if n in r[ki]][4:])
PY
python tools/knn_model_stats.py 2>&1 | tail -4
python tools/profile_calls.py pseg 16 2048 40 2>&1 | grep -E "knn|sum"
