python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
python tools/fp_calls.py 2>&1 | tail -40
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; python -c "
import json;d=json.load(open('gpurun_out/bench_f.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['roofline']['eager_ms_per_step']);print(d['extra']['cfg3']['clouds_per_s'],d['extra']['cfg4']['clouds_per_s'])"
