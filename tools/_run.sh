python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; tail -2 gpurun_out/t_all.log
python - <<PY
import contextlib, io, os, sys, torch
sys.path.insert(0,'.')
import svnet_b200 as sv
from svnet_b200.synthetic import make_args, one_hot_labels, synthetic_clouds, synthetic_state_dict
with contextlib.redirect_stdout(io.StringIO()):
    net = sv.SV_DGCNN_PSEG(make_args(k=40, binary=True), 50)
net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1004)); net = net.cuda().eval()
for B in (16, 128):
    x = synthetic_clouds(B, 2048, 1004).cuda(); l = one_hot_labels(B).cuda()
    with torch.no_grad():
        g = sv.GraphedForward(net, x, l)
        for _ in range(3): g(x, l)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g(x, l); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        print("pseg graph B=%d: %.3f ms  %.0f clouds/s" % (B, best, B / best * 1e3))
PY
