timeout 120 python tools/fp_tc_check.py 32 > gpurun_out/fp_tc.log 2>&1; echo rc=$?; tail -3 gpurun_out/fp_tc.log
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
python tools/profile_calls.py cls_fp 32 1024 20 2>&1 | tail -14
