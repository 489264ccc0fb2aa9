python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
python tools/pseg_calls.py 16 2>&1 | grep -E "whole|seg_head|pool|sum"
python tools/profile_calls.py 2>&1 | grep -E "svfuse|sum"
