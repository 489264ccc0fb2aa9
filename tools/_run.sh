timeout 120 python tools/fp_tc_check.py 32 > gpurun_out/fp_tc.log 2>&1; echo rc=$?; tail -8 gpurun_out/fp_tc.log
python -m pytest tests -m gpu -x -q -k "reference or model_c_entry" 2>&1 | tail -1
