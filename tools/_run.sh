python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('bench:',round(d['value']),d['ms_per_step'],round(d['e2e']['value']));e=d['extra'];print(e['cfg3']['clouds_per_s'],e['cfg4']['clouds_per_s'],[c['clouds_per_s'] for c in e['cfg5']])"
