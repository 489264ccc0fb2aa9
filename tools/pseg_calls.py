"""Developer tool: per-call CUDA-event times, in call order, of one part-segmentation forward (cfg4 shape)."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svnet_b200 as sv
from svnet_b200 import _native as nv, fused
from svnet_b200.synthetic import make_args, one_hot_labels, synthetic_clouds, synthetic_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
with contextlib.redirect_stdout(io.StringIO()):
    net = sv.SV_DGCNN_PSEG(make_args(k=40, binary=True), 50)
net.load_state_dict(synthetic_state_dict(net.state_dict(), seed=1004))
net = net.cuda().eval()
x = synthetic_clouds(B, 2048, 1004).cuda()
l = one_hot_labels(B).cuda()
fused.CONCURRENT_HALVES = False
with torch.no_grad():
    for _ in range(3): net(x, l)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net(x, l); e1.record(); torch.cuda.synchronize()
    print("whole forward (eager, incl. torch glue) %.3f ms" % e0.elapsed_time(e1))
    nv.PROFILE[0] = set(nv.EXPORTS); nv.ORDER.clear()
    net(x, l); torch.cuda.synchronize()
tot = 0.0
for name, e0, e1 in nv.ORDER:
    t = e0.elapsed_time(e1); tot += t
    print("%-28s %8.1f us" % (name, 1e3 * t))
print("sum %.3f ms for %d clouds" % (tot, B))
