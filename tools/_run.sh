python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "model_c_entry" > gpurun_out/t_model.log 2>&1; tail -25 gpurun_out/t_model.log
