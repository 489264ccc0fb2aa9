SECONDS=0
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo rc=$? secs=$SECONDS; tail -3 gpurun_out/bench_g.err; python -c "
import json;d=json.load(open('gpurun_out/bench_g.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['roofline']['frac']);print(json.dumps(d['extra'],indent=0)[:3000])"
