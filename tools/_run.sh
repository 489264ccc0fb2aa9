ncu --set full --clock-control none -k regex:"edge_fp_tc_kernel" --launch-skip 9 --launch-count 3 -o gpurun_out/r2i_fp python tools/fp_calls.py > gpurun_out/ncu_fp.log 2>&1
ncu --set full --clock-control none -k regex:"signpack|knn_finish_wide|seg_transpose|pool_rows" --launch-skip 30 --launch-count 12 -o gpurun_out/r2i_seg python tools/pseg_calls.py 16 > gpurun_out/ncu_seg.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2i_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel'],d['roofline']['kernel_ms'],d['roofline']['frac']);e=d['extra'];print(e['cfg3']['clouds_per_s'],e['cfg4']['clouds_per_s'],[c['clouds_per_s'] for c in e['cfg5']])"
ls -la gpurun_out/r2i*
