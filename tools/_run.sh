python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "model_c_entry" > gpurun_out/t_model.log 2>&1; tail -15 gpurun_out/t_model.log
python - <<PY
import contextlib, io, sys, torch
sys.path.insert(0,'.')
import svnet_b200 as sv
from svnet_b200.synthetic import make_args, synthetic_clouds, synthetic_state_dict
for binary, B in ((True, 32), (False, 32), (False, 256)):
    with contextlib.redirect_stdout(io.StringIO()):
        net = sv.SV_DGCNN_CLS(make_args(k=20, binary=binary), 40)
    sd = synthetic_state_dict(net.state_dict(), seed=1002)
    native = sv.NativeModel("SV_DGCNN_CLS", sd, k=20, binary=binary, num_class=40)
    x = synthetic_clouds(B, 1024, 1002).cuda()
    for _ in range(3): native(x)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); native(x); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    print("svnet_model_forward binary=%s B=%d: %.3f ms  %.0f clouds/s (eager, one stream)" % (binary, B, best, B / best * 1e3))
PY
