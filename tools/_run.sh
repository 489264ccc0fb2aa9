python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; python -c "
import json;d=json.load(open('gpurun_out/bench_b.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['roofline']['eager_ms_per_step']);print(d['extra'])"
