set -x
python tools/one_forward.py > gpurun_out/one_fwd.log 2>&1
ncu --set full --clock-control none -o gpurun_out/r2h_all python tools/one_forward.py > gpurun_out/ncu_all.log 2>&1
ncu -i gpurun_out/r2h_all.ncu-rep --page raw --csv > gpurun_out/r2h_raw.csv 2>/dev/null
rm -f gpurun_out/r2h_all.ncu-rep
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python tools/ref_parity.py r2h > gpurun_out/refparity_h.log 2>&1; tail -3 gpurun_out/refparity_h.log | cut -c1-400
python tools/sweep.py > gpurun_out/r2h_sweep_configs.json 2> gpurun_out/sweep.err; tail -2 gpurun_out/sweep.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2h_bench_reference_arm.json 2> gpurun_out/r2h_ref.err
ls -la gpurun_out | tail -12
