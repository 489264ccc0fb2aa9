python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gate_|head_" --csv --log-file gpurun_out/l.csv python tools/one_forward.py > gpurun_out/ncu_a.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/l.csv') if not l.startswith('=='))]
h=rows[0];vi=h.index("Metric Value");ki=h.index("Kernel Name")
print([(r[ki][:30],float(r[vi])/1e3) for r in rows[1:]][6:])
PY
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; python -c "
import json;d=json.load(open('gpurun_out/bench_d.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['roofline']['eager_ms_per_step']);print(d['extra']['cfg3']['clouds_per_s'],d['extra']['cfg4']['clouds_per_s'])"
